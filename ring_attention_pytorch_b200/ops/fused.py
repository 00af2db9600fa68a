"""Thin Python wrappers over the sm_100a kernels (``torch.ops.rab.*``).

These are the building blocks the autograd ops in :mod:`ring_attention_pytorch_b200.ops.ring_cuda`
compose; they are also what the GPU unit tests drive directly.  ``emulate_ring_forward`` runs a whole
W-rank ring on ONE device by giving every emulated rank its own K/V gather buffer and pointing the
"peer" addresses at the other ranks' buffers – the kernel cannot tell the difference, which lets the
multi-hop fetch/ready-flag protocol be tested on a single GPU.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from ring_attention_pytorch_b200.ops import _ext
from ring_attention_pytorch_b200.parallel.layout import (PositionMap, make_position_map, ring_hop_owners,
                                                          ring_query_owners)


def pack_key_mask_bits(mask: torch.Tensor) -> torch.Tensor:
    """[world, b, n] bool (True = keep) -> [world, b, words] int32 bit-packed, words % 4 == 0."""
    world, b, n = mask.shape
    words = ((n + 127) // 128) * 4
    padded = torch.zeros(world, b, words * 32, dtype=torch.bool, device=mask.device)
    padded[..., :n] = mask
    bits = padded.view(world, b, words, 32).to(torch.int64)
    weights = (1 << torch.arange(32, device=mask.device, dtype=torch.int64))
    packed = (bits * weights).sum(-1)
    packed = torch.where(packed >= 2 ** 31, packed - 2 ** 32, packed)
    return packed.to(torch.int32).contiguous()


def alloc_kv_buffer(world: int, batch: int, kv_heads: int, n_k: int, d: int, dtype, device) -> torch.Tensor:
    return torch.empty(world, 2, batch * kv_heads, n_k, d, dtype=dtype, device=device)


def fused_attn_fwd(
    q: torch.Tensor,
    kv_buf: torch.Tensor,
    peer_ptrs: Sequence[int],
    ready: torch.Tensor,
    kmask_bits: Optional[torch.Tensor],
    *,
    kv_heads: int,
    rank: int,
    pm: PositionMap,
    causal: bool,
    window: Optional[int],
    scale: float,
    softclamp: float = 0.0,
    q_pos_offset: int = 0,
    hop_owner: Optional[List[int]] = None,
):
    if hop_owner is None:
        hop_owner = ring_hop_owners(pm, rank, causal, window)
    return _ext.ops().attn_fwd(
        q, kv_buf, list(peer_ptrs), ready, kmask_bits, kv_heads, rank, bool(causal), int(window or 0), float(scale),
        float(softclamp), pm.stride, pm.seg_len, pm.base0, pm.base1, int(q_pos_offset), list(hop_owner))


def emulate_ring_forward(
    qs: Sequence[torch.Tensor],
    ks: Sequence[torch.Tensor],
    vs: Sequence[torch.Tensor],
    *,
    layout: str = "plain",
    causal: bool = False,
    window: Optional[int] = None,
    softclamp: float = 0.0,
    key_masks: Optional[Sequence[torch.Tensor]] = None,
    scale: Optional[float] = None,
):
    """Run the fused forward for every rank of a W-rank ring on the current device.

    qs/ks/vs: per-rank shards ``[b, n, h, d]`` / ``[b, n, hk, d]``.  Returns (outs, lses) lists.
    """
    ops = _ext.ops()
    world = len(qs)
    b, n, h, d = qs[0].shape
    hk = ks[0].shape[2]
    dev, dt = qs[0].device, qs[0].dtype
    pm = make_position_map(layout, world, n)
    scale = d ** -0.5 if scale is None else scale
    bufs = [alloc_kv_buffer(world, b, hk, n, d, dt, dev) for _ in range(world)]
    for r in range(world):
        bufs[r].zero_()
        ops.pack_kv(ks[r], vs[r], bufs[r][r])
    kbits = None
    if key_masks is not None:
        kbits = pack_key_mask_bits(torch.stack(list(key_masks), 0))
    outs, lses = [], []
    for r in range(world):
        ready = torch.zeros(world, dtype=torch.int32, device=dev)
        peers = [bufs[o][o].data_ptr() for o in range(world)]  # owner o's own slot
        o, lse = fused_attn_fwd(qs[r].contiguous(), bufs[r], peers, ready, kbits, kv_heads=hk, rank=r, pm=pm,
                                causal=causal, window=window, scale=scale, softclamp=softclamp)
        outs.append(o)
        lses.append(lse)
    return outs, lses


def pad64(n: int) -> int:
    return (n + 63) // 64 * 64


def alloc_qdo_buffer(world: int, batch: int, heads: int, n_q: int, d: int, dtype, device) -> torch.Tensor:
    return torch.empty(world, 2, batch * heads, n_q, d, dtype=dtype, device=device)


def alloc_stat_buffer(world: int, batch: int, heads: int, n_q: int, device) -> torch.Tensor:
    return torch.zeros(world, 2, batch * heads, pad64(n_q), dtype=torch.float32, device=device)


def fused_attn_bwd(
    qdo_buf: torch.Tensor,
    kv_buf: torch.Tensor,
    stat_buf: torch.Tensor,
    kmask_bits: Optional[torch.Tensor],
    *,
    batch: int,
    heads: int,
    kv_heads: int,
    rank: int,
    pm: PositionMap,
    causal: bool,
    window: Optional[int],
    scale: float,
    softclamp: float = 0.0,
    q_pos_offset: int = 0,
    ready_kv: Optional[torch.Tensor] = None,
    ready_q: Optional[torch.Tensor] = None,
    ready_target: int = 0,
):
    """Run both backward kernels on already gathered buffers; returns (dq, dk, dv)."""
    ops = _ext.ops()
    kv_owners = ring_hop_owners(pm, rank, causal, window)
    q_owners = ring_query_owners(pm, rank, causal, window)
    common = (kmask_bits, batch, heads, kv_heads, rank, bool(causal), int(window or 0), float(scale), float(softclamp),
              pm.stride, pm.seg_len, pm.base0, pm.base1, int(q_pos_offset))
    dq = ops.attn_bwd_dq(qdo_buf, kv_buf, stat_buf, ready_kv, ready_target, *common, kv_owners)
    dk, dv = ops.attn_bwd_dkdv(qdo_buf, kv_buf, stat_buf, ready_q, ready_target, *common, q_owners)
    return dq, dk, dv


def fused_attn_bwd_one_kernel(
    qdo_buf: torch.Tensor,
    kv_buf: torch.Tensor,
    stat_buf: torch.Tensor,
    kmask_bits: Optional[torch.Tensor],
    *,
    batch: int,
    heads: int,
    kv_heads: int,
    n_q: int,
    pm: PositionMap,
    causal: bool,
    window: Optional[int],
    scale: float,
    softclamp: float = 0.0,
    q_pos_offset: int = 0,
):
    """EXPERIMENTAL (compile-checked, not yet validated on a GPU): the whole backward in the KV-stationary kernel — S and
    dP are computed once (5 GEMMs instead of 7), dQ^T = K^T dS^T is reduced into an fp32 accumulator with
    ``red.global.add.f32``.  Single rank, head dim 128 only.  Returns (dq fp32 [b, n_q, h, d], dk, dv)."""
    ops = _ext.ops()
    d = kv_buf.shape[-1]
    dq_acc = torch.zeros(batch, n_q, heads, d, dtype=torch.float32, device=kv_buf.device)
    dk, dv = ops.attn_bwd_fused(qdo_buf, kv_buf, stat_buf, None, 0, kmask_bits, batch, heads, kv_heads, 0, bool(causal),
                                int(window or 0), float(scale), float(softclamp), pm.stride, pm.seg_len, pm.base0,
                                pm.base1, int(q_pos_offset), [0], dq_acc)
    return dq_acc, dk, dv


def emulate_ring_backward(
    qs, ks, vs, outs, lses, douts, *,
    layout: str = "plain",
    causal: bool = False,
    window: Optional[int] = None,
    softclamp: float = 0.0,
    key_masks=None,
    scale: Optional[float] = None,
):
    """Backward of :func:`emulate_ring_forward` for every emulated rank on the current device."""
    ops = _ext.ops()
    world = len(qs)
    b, n, h, d = qs[0].shape
    hk = ks[0].shape[2]
    dev, dt = qs[0].device, qs[0].dtype
    pm = make_position_map(layout, world, n)
    scale = d ** -0.5 if scale is None else scale
    kv_all = alloc_kv_buffer(world, b, hk, n, d, dt, dev)
    qdo_all = alloc_qdo_buffer(world, b, h, n, d, dt, dev)
    stat_all = alloc_stat_buffer(world, b, h, n, dev)
    for r in range(world):
        ops.pack_kv(ks[r], vs[r], kv_all[r])
        ops.bwd_prep(qs[r].contiguous(), outs[r].contiguous(), douts[r].contiguous(), lses[r].contiguous(), qdo_all,
                     stat_all, r)
    kbits = None
    if key_masks is not None:
        kbits = pack_key_mask_bits(torch.stack(list(key_masks), 0))
    res = []
    for r in range(world):
        # every emulated rank sees the same fully gathered buffers (what the NVLink gather produces)
        res.append(fused_attn_bwd(qdo_all, kv_all, stat_all, kbits, batch=b, heads=h, kv_heads=hk, rank=r, pm=pm,
                                  causal=causal, window=window, scale=scale, softclamp=softclamp))
    return res
