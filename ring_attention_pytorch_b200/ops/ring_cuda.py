"""Ring flash attention on the sm_100a kernels (autograd Function + reference-compatible wrapper).

Public surface mirrors reference ring_flash_attention_cuda.py:353-371 (``ring_flash_attn_cuda`` /
``ring_flash_attn_cuda_``) plus a ``layout`` argument ('plain' | 'striped' | 'zigzag').

Forward, per rank (one stream, no host synchronisation, no NCCL on the hot path):

    pack_kv (K,V -> head-major slot) -> copy to symmetric staging -> device barrier ->
    ONE fused kernel: tcgen05 flash attention over every hop of the ring while its fetcher warps pull
    the other ranks' K/V slots over NVLink (bulk TMA) into the local gather buffer

Backward: ``bwd_prep`` (delta, lse->log2, Q/dO head-major) -> staging -> barrier -> copy engines pull the
peers' Q/dO/stat slots on a side stream *while* the dQ kernel (which only needs the K/V gather saved by
the forward) runs -> dK/dV kernel.  Every rank finishes its own dQ, dK, dV: no reduction, no atomics.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import Tensor
from torch.autograd import Function

from ring_attention_pytorch_b200.ops import _ext
from ring_attention_pytorch_b200.ops.fused import (
    alloc_kv_buffer,
    alloc_qdo_buffer,
    alloc_stat_buffer,
    fused_attn_bwd,
    fused_attn_fwd,
    pack_key_mask_bits,
)
from ring_attention_pytorch_b200.parallel.distributed import default, exists, get_rank, get_world_size, is_distributed
from ring_attention_pytorch_b200.parallel.layout import make_position_map, ring_hop_owners, ring_query_owners
from ring_attention_pytorch_b200.parallel.symm import get_workspace

# counts launches of our own kernels (bench.py reports it as gpu_launches)
LAUNCHES = {"count": 0}

# save_kv_gather=True : the forward's gathered K/V ([W, 2, b*hk, n, d], i.e. the whole ring's K/V) is kept for the
#                       backward, whose dQ kernel then needs no communication at all (fastest).
# save_kv_gather=False: only this rank's K/V slot is kept (O(n) activation memory per layer, the classic ring
#                       attention footprint); the backward re-pulls the peers' slots with the copy engines before
#                       the dQ kernel starts.
# fused_backward=True  : EXPERIMENTAL, not yet validated on a GPU — single-rank, head-dim-128 calls run the whole
#                        backward in one KV-stationary kernel (5 GEMMs, dQ through fp32 global reductions).
CONFIG = {"save_kv_gather": True, "fused_backward": False}


def _count(n: int = 1) -> None:
    LAUNCHES["count"] += n


def _pad_head_dim(t: Tensor, d_pad: int) -> Tensor:
    d = t.shape[-1]
    return t if d == d_pad else F.pad(t, (0, d_pad - d))


def _gather_ring_masks(mask: Tensor, ring_size: int) -> Tensor:
    """[b, n] bool on every rank -> [ring, b, n] for this rank's ring set (cold path, NCCL)."""
    world = get_world_size()
    gathered = [torch.empty_like(mask, dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(gathered, mask.to(torch.uint8).contiguous())
    ring_set = get_rank() // ring_size
    return torch.stack(gathered[ring_set * ring_size:(ring_set + 1) * ring_size]).bool()


class RingFlashAttentionCUDAFunction(Function):
    @staticmethod
    def forward(
        ctx,
        q: Tensor,
        k: Tensor,
        v: Tensor,
        mask: Optional[Tensor],
        causal: bool,
        bucket_size: int,
        ring_reduce_col: bool,
        striped_ring_attn: bool,
        max_lookback_seq_len: Optional[int],
        ring_size: Optional[int],
        softclamp_qk_sim: bool = False,
        softclamp_value: float = 50.0,
        layout: Optional[str] = None,
    ):
        assert q.is_cuda and k.is_cuda and v.is_cuda, "ring_flash_attn_cuda needs CUDA tensors"
        ops = _ext.ops()
        orig_dtype = q.dtype
        if q.dtype not in (torch.bfloat16, torch.float16):
            q, k, v = (t.to(torch.bfloat16) for t in (q, k, v))  # reference casts fp32 to fp16 (_cuda.py:72-79)
        dt = q.dtype
        k, v = k.to(dt), v.to(dt)

        ring_size = default(ring_size, get_world_size())
        cross_attn = q.shape[1] != k.shape[1]
        use_ring = bool(ring_reduce_col) and is_distributed() and not cross_attn and ring_size > 1
        layout = default(layout, "striped" if (striped_ring_attn and use_ring) else "plain")
        if not use_ring:
            layout, ring_size = "plain", 1
        assert not (exists(max_lookback_seq_len) and not causal), "look-back windows need causal attention"
        if causal:
            mask = None  # reference _cuda.py:105-106

        b, n_q, h, d = q.shape
        n_k, hk = k.shape[1], k.shape[2]
        assert h % hk == 0 and v.shape == k.shape
        assert d <= 128, "head dimension up to 128 is supported"
        d_pad = 64 if d <= 64 else 128
        scale = d ** -0.5
        qp, kp, vp = (_pad_head_dim(t, d_pad).contiguous() for t in (q, k, v))

        rank = get_rank() % ring_size if use_ring else 0
        pm = make_position_map(layout, ring_size, n_k)
        q_off = (n_k - n_q) if (cross_attn and causal) else 0
        dev = q.device

        kv_gather = alloc_kv_buffer(ring_size, b, hk, n_k, d_pad, dt, dev)
        ops.pack_kv(kp, vp, kv_gather[rank])
        _count()
        ready = torch.zeros(ring_size, dtype=torch.int32, device=dev)
        peers = [0] * ring_size
        kbits = None
        if use_ring:
            ws = get_workspace(ring_size, dev)
            slot_bytes = kv_gather[rank].numel() * kv_gather.element_size()
            stage, peer_ptrs = ws.staging(f"kv", slot_bytes)
            stage.copy_(kv_gather[rank].view(torch.uint8).reshape(-1))
            ws.barrier()
            _count(2)
            peers = [0 if o == rank else peer_ptrs[o] for o in range(ring_size)]
            if exists(mask):
                kbits = pack_key_mask_bits(_gather_ring_masks(mask, ring_size))
        elif exists(mask):
            kbits = pack_key_mask_bits(mask[None])

        softclamp = float(softclamp_value) if softclamp_qk_sim else 0.0
        o, lse = fused_attn_fwd(qp, kv_gather, peers, ready, kbits, kv_heads=hk, rank=rank, pm=pm, causal=causal,
                                window=max_lookback_seq_len, scale=scale, softclamp=softclamp, q_pos_offset=q_off)
        _count()

        keep_gather = CONFIG["save_kv_gather"] or not use_ring
        ctx.cfg = (causal, max_lookback_seq_len, ring_size, rank, layout, softclamp, scale, q_off, use_ring, d, d_pad,
                   orig_dtype, hk, keep_gather)
        kv_saved = kv_gather if keep_gather else kv_gather[rank].clone()
        ctx.save_for_backward(qp, o, lse, kv_saved, kbits if kbits is not None else torch.empty(0, device=dev))
        out = o[..., :d]
        return out.to(orig_dtype) if orig_dtype != dt else out

    @staticmethod
    def backward(ctx, do: Tensor):
        ops = _ext.ops()
        (causal, window, ring_size, rank, layout, softclamp, scale, q_off, use_ring, d, d_pad, orig_dtype,
         hk, keep_gather) = ctx.cfg
        qp, o, lse, kv_saved, kbits = ctx.saved_tensors
        kbits = kbits if kbits.numel() > 0 else None
        dt = qp.dtype
        b, n_q, h, _ = qp.shape
        dev = qp.device
        if keep_gather:
            kv_gather = kv_saved
        else:  # memory-lean mode: rebuild the gather buffer around this rank's own slot
            kv_gather = alloc_kv_buffer(ring_size, b, hk, kv_saved.shape[2], d_pad, dt, dev)
            kv_gather[rank].copy_(kv_saved)
        n_k = kv_gather.shape[3]
        pm = make_position_map(layout, ring_size, n_k)
        dop = _pad_head_dim(do.to(dt), d_pad).contiguous()

        qdo_gather = alloc_qdo_buffer(ring_size, b, h, n_q, d_pad, dt, dev)
        stat_gather = alloc_stat_buffer(ring_size, b, h, n_q, dev)
        ops.bwd_prep(qp, o, dop, lse, qdo_gather, stat_gather, rank)
        _count()

        gather_done = None
        kv_done = None
        if use_ring:
            ws = get_workspace(ring_size, dev)
            qdo_bytes = qdo_gather[rank].numel() * qdo_gather.element_size()
            stat_bytes = stat_gather[rank].numel() * 4
            stage, peer_ptrs = ws.staging("qdo", qdo_bytes + stat_bytes)
            stage[:qdo_bytes].copy_(qdo_gather[rank].view(torch.uint8).reshape(-1))
            stage[qdo_bytes:qdo_bytes + stat_bytes].copy_(stat_gather[rank].view(torch.uint8).reshape(-1))
            kv_peer_ptrs = None
            if not keep_gather:
                kv_bytes = kv_gather[rank].numel() * kv_gather.element_size()
                kv_stage, kv_peer_ptrs = ws.staging("kv", kv_bytes)
                kv_stage.copy_(kv_gather[rank].view(torch.uint8).reshape(-1))
            ws.barrier()
            _count(3)
            main = torch.cuda.current_stream(dev)
            start = torch.cuda.Event()
            start.record(main)
            q_owners = ring_query_owners(pm, rank, causal, window)
            with torch.cuda.stream(ws.side_stream):
                ws.side_stream.wait_event(start)
                if kv_peer_ptrs is not None:  # K/V first: the dQ kernel is waiting for it
                    for o_rank in ring_hop_owners(pm, rank, causal, window)[1:]:
                        ops.peer_copy(kv_gather[o_rank], kv_peer_ptrs[o_rank], kv_bytes)
                    kv_done = torch.cuda.Event()
                    kv_done.record(ws.side_stream)
                    kv_gather.record_stream(ws.side_stream)
                for o_rank in q_owners[1:]:
                    ops.peer_copy(qdo_gather[o_rank], peer_ptrs[o_rank], qdo_bytes)
                    ops.peer_copy(stat_gather[o_rank], peer_ptrs[o_rank] + qdo_bytes, stat_bytes)
                gather_done = torch.cuda.Event()
                gather_done.record(ws.side_stream)
            qdo_gather.record_stream(ws.side_stream)
            stat_gather.record_stream(ws.side_stream)

        if CONFIG["fused_backward"] and not use_ring and d_pad == 128:
            from ring_attention_pytorch_b200.ops.fused import fused_attn_bwd_one_kernel

            dq32, dk, dv = fused_attn_bwd_one_kernel(qdo_gather, kv_gather, stat_gather, kbits, batch=b, heads=h,
                                                     kv_heads=hk, n_q=n_q, pm=pm, causal=causal, window=window,
                                                     scale=scale, softclamp=softclamp, q_pos_offset=q_off)
            _count(1)
            dq = dq32.to(dt)
            dq, dk, dv = dq[..., :d], dk[..., :d], dv[..., :d]
            if orig_dtype != dt:
                dq, dk, dv = dq.to(orig_dtype), dk.to(orig_dtype), dv.to(orig_dtype)
            return dq, dk, dv, None, None, None, None, None, None, None, None, None, None

        common = (kbits, b, h, hk, rank, bool(causal), int(window or 0), float(scale), float(softclamp), pm.stride,
                  pm.seg_len, pm.base0, pm.base1, int(q_off))
        # dQ only needs the K/V gather (saved by the forward, or just re-pulled): it overlaps with the Q/dO gather
        if kv_done is not None:
            torch.cuda.current_stream(dev).wait_event(kv_done)
        dq = ops.attn_bwd_dq(qdo_gather, kv_gather, stat_gather, None, 0, *common,
                             ring_hop_owners(pm, rank, causal, window))
        if gather_done is not None:
            torch.cuda.current_stream(dev).wait_event(gather_done)
        dk, dv = ops.attn_bwd_dkdv(qdo_gather, kv_gather, stat_gather, None, 0, *common,
                                   ring_query_owners(pm, rank, causal, window))
        _count(2)

        dq, dk, dv = dq[..., :d], dk[..., :d], dv[..., :d]
        if orig_dtype != dt:
            dq, dk, dv = dq.to(orig_dtype), dk.to(orig_dtype), dv.to(orig_dtype)
        return dq, dk, dv, None, None, None, None, None, None, None, None, None, None


ring_flash_attn_cuda_ = RingFlashAttentionCUDAFunction.apply


@torch.autocast("cuda", enabled=False)
def ring_flash_attn_cuda(
    q: Tensor,
    k: Tensor,
    v: Tensor,
    mask: Optional[Tensor] = None,
    causal: bool = False,
    bucket_size: int = 1024,
    ring_reduce_col: bool = False,
    striped_ring_attn: bool = False,
    max_lookback_seq_len: Optional[int] = None,
    ring_size: Optional[int] = None,
    softclamp_qk_sim: bool = False,
    softclamp_value: float = 50.0,
    layout: Optional[str] = None,
) -> Tensor:
    """q [b, n, h, d]; k, v [b, n, hk, d] (this rank's shard when ``ring_reduce_col``).  ``bucket_size`` is
    accepted for signature parity; tiling is fixed by the kernel (128 x 128)."""
    return ring_flash_attn_cuda_(q, k, v, mask, causal, bucket_size, ring_reduce_col, striped_ring_attn,
                                 max_lookback_seq_len, ring_size, softclamp_qk_sim, softclamp_value, layout)
