"""Zig-zag context parallelism (Llama-3 style; reference zig_zag_attention.py:35-140).

``zig_zag_pad_seq`` / ``zig_zag_shard`` keep the reference's interface.  ``zig_zag_attn`` has two modes:

* ``causal=True`` (new, preferred): the zig-zag layout is just another position map of the ring kernels
  – K/V are pulled tile-by-tile over NVLink inside the fused kernel (CUDA) or ride the P2P ring (portable
  path); nothing is all-gathered and no mask is materialised.  The reference all-gathers K and V and needs
  a caller-built dense ``[n_local, N]`` boolean mask (zig_zag_attention.py:123-138), which is O(N) memory
  per rank for K/V and ~137 GB of mask per rank at N = 1M, W = 8.
* ``attn_mask=<dense bool mask>``: reference-compatible behaviour (all-gather + SDPA) for arbitrary masks.
"""
from __future__ import annotations

from collections import namedtuple
from math import ceil
from typing import Optional

import torch
import torch.nn.functional as F
from torch import Tensor

from ring_attention_pytorch_b200.parallel.distributed import AllGather, get_rank, get_world_size, is_distributed
from ring_attention_pytorch_b200.parallel.layout import make_position_map
from ring_attention_pytorch_b200.utils.validate import typecheck

ShardOutput = namedtuple("ShardOutput", ["local_sequence", "query_positions", "key_value_positions"])


def zig_zag_pad_seq(t: Tensor):
    """Pad the sequence (dim -2) to a multiple of ``2 * world`` (reference zig_zag_attention.py:35-45)."""
    seq_len = t.shape[-2]
    chunks = 2 * get_world_size()
    padded = ceil(seq_len / chunks) * chunks
    t = F.pad(t, (0, 0, 0, padded - seq_len), value=0.0)

    def inverse(out: Tensor) -> Tensor:
        return out[..., :seq_len, :]

    return t, inverse


def zig_zag_shard(t: Tensor, all_gather_batch: bool = False):
    """Rank r keeps chunks ``r`` and ``2W-1-r`` of ``2W`` (reference zig_zag_attention.py:55-100)."""
    rank, world = get_rank(), get_world_size()
    gather_sizes = None
    if all_gather_batch:
        t, gather_sizes = AllGather(dim=0)(t)
    seq_len = t.shape[-2]
    assert seq_len % (2 * world) == 0, "pad with zig_zag_pad_seq first"
    n_local = seq_len // world
    pm = make_position_map("zigzag", world, n_local)
    q_indices = pm.positions(rank, t.device)
    kv_indices = torch.cat([pm.positions(r, t.device) for r in range(world)])
    local = t.index_select(-2, q_indices).contiguous()

    def inverse(two_chunks: Tensor) -> Tensor:
        gathered, _ = AllGather(dim=-2)(two_chunks)
        inv = torch.empty_like(kv_indices)
        inv[kv_indices] = torch.arange(seq_len, device=kv_indices.device)
        out = gathered.index_select(-2, inv)
        if all_gather_batch:
            out = out.split(gather_sizes.tolist(), dim=0)[rank]
        return out

    return ShardOutput(local, q_indices, kv_indices), inverse


@typecheck
def zig_zag_attn(
    q: Tensor,
    k: Tensor,
    v: Tensor,
    dropout: float = 0.0,
    attn_mask: Optional[Tensor] = None,
    causal: Optional[bool] = None,
) -> Tensor:
    """q [b, qh, i, d]; k, v [b, h, j, d] (this rank's zig-zag shard).  Returns [b, qh, i, d]."""
    heads, kv_heads = q.shape[1], k.shape[1]
    assert heads % kv_heads == 0
    if causal is None:
        causal = attn_mask is None and False

    if attn_mask is None and causal and dropout > 0.0:
        # The fused ring kernels have no dropout.  Honour the argument like the reference does (it forwards dropout to
        # SDPA, zig_zag_attention.py:134-138) by taking the gathered dense path with the causal mask of the zig-zag
        # positions instead of silently ignoring it.
        world = get_world_size() if is_distributed() else 1
        rank = get_rank() if is_distributed() else 0
        pm = make_position_map("zigzag" if world > 1 else "plain", world, q.shape[-2])
        q_pos = pm.positions(rank, q.device)
        k_pos = torch.cat([pm.positions(r, q.device) for r in range(world)])
        attn_mask = q_pos[:, None] >= k_pos[None, :]

    if attn_mask is None and causal:
        # ring schedule with the zig-zag position map: no all-gather, no dense mask
        qn, kn, vn = (t.transpose(1, 2) for t in (q, k, v))
        if q.is_cuda:
            from ring_attention_pytorch_b200.ops.ring_cuda import ring_flash_attn_cuda

            out = ring_flash_attn_cuda(qn, kn, vn, None, True, 1024, True, False, None, None, False, 50.0, "zigzag")
        else:
            from ring_attention_pytorch_b200.ops.ring_flash_naive import ring_flash_attn

            out = ring_flash_attn(qn, kn, vn, None, True, 1024, True, False, None, None, False, 50.0, "zigzag")
        return out.transpose(1, 2)

    if is_distributed():
        gather = AllGather(dim=-2)
        k, _ = gather(k)
        v, _ = gather(v)
    g = heads // kv_heads
    k, v = (t.repeat(1, g, 1, 1) for t in (k, v))  # query head j <- kv head j % kv_heads
    return F.scaled_dot_product_attention(q, k, v, dropout_p=dropout, attn_mask=attn_mask)
