"""Tree-attention decoding: single-query attention over a KV cache sharded along the sequence across ranks
(reference tree_attn_decoding.py:23-104, Algorithm 3 of https://arxiv.org/abs/2408.04093).

Layout is head-first like the reference: ``q [b, h, 1, d]``, ``k [b, hk, n, d]``, ``v [b, hk, n, dv]``.

* CUDA path: ``csrc/tree_decode_sm100.cu`` – ONE persistent cooperative kernel per rank and step computes the
  split-KV partials of its KV shard, merges the splits, publishes ``(out, lse)`` in symmetric memory, signals the
  peers and merges all ranks' partials in-kernel (NVLink peer loads, or ``multimem.ld_reduce`` through the NVSwitch
  when the buffers have a multicast mapping), replacing the reference's Triton launch padded to a 128-row tile plus
  three latency-bound all-reduces (MAX, SUM, SUM).
* portable path (CPU / gloo): local einsum attention + one MAX and one packed SUM all-reduce.

Fixes vs. the reference: ``shard_kv_seq=False`` with ``k=None`` works (reference uses an undefined ``dim_v``
– tree_attn_decoding.py:46 vs 84) by taking ``dim_v`` explicitly; grouped-query heads are supported.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist
from torch import Tensor

from ring_attention_pytorch_b200.parallel.distributed import default, exists, get_rank, get_world_size, is_distributed
from ring_attention_pytorch_b200.utils.validate import typecheck


def _local_attention(q: Tensor, k: Tensor, v: Tensor):
    """q [b,h,1,d], k [b,hk,n,d], v [b,hk,n,dv] -> (out [b,h,1,dv] fp32, lse [b,h,1,1] fp32)."""
    b, h, _, d = q.shape
    hk = k.shape[1]
    g = h // hk
    scale = d ** -0.5
    qf = q.float().view(b, g, hk, 1, d)  # query head j uses kv head j % hk
    sim = torch.einsum("bghid,bhjd->bghij", qf, k.float()) * scale
    lse = sim.logsumexp(dim=-1, keepdim=True)
    attn = (sim - lse).exp()
    out = torch.einsum("bghij,bhjd->bghid", attn, v.float())
    return out.reshape(b, h, 1, -1), lse.reshape(b, h, 1, 1)


@torch.no_grad()
@typecheck
def tree_attn_decode(
    q: Tensor,
    k: Optional[Tensor] = None,
    v: Optional[Tensor] = None,
    eps: float = 1e-8,
    shard_kv_seq: bool = True,
    use_triton: Optional[bool] = None,
    dim_v: Optional[int] = None,
) -> Tensor:
    """Returns ``[b, h, 1, dv]`` in ``q.dtype``.

    ``shard_kv_seq=True``: every rank passes the *full* K/V and attends to its own ``chunk(world)`` slice
    (ranks beyond the number of chunks contribute nothing).  ``shard_kv_seq=False``: K/V are already this
    rank's shard (``None`` for an empty shard).  ``use_triton`` is kept for signature parity and selects the
    sm_100a kernel (default: on CUDA inputs).
    """
    assert not (exists(k) ^ exists(v)), "keys and values are either both None, or both present"
    dtype = q.dtype
    b, h = q.shape[:2]
    if exists(v):
        dim_v = v.shape[-1]

    if shard_kv_seq:
        assert exists(k), "keys and values must be passed if not already sharded across sequence"
        rank, world = get_rank(), get_world_size()
        ks, vs = k.chunk(world, dim=-2), v.chunk(world, dim=-2)
        k, v = (ks[rank], vs[rank]) if rank < len(ks) else (None, None)
    assert exists(dim_v), "dim_v is required when this rank holds no keys"

    use_kernel = default(use_triton, q.is_cuda)
    assert not (use_kernel and not q.is_cuda), "input needs to be on cuda to use the sm_100a kernel"

    if use_kernel:
        from ring_attention_pytorch_b200.ops.tree_decode_cuda import tree_decode_cuda

        return tree_decode_cuda(q, k, v, dim_v=dim_v, eps=eps).to(dtype)

    if exists(k) and k.shape[-2] > 0:
        local_out, lse = _local_attention(q, k, v)
    else:
        local_out = q.new_zeros((b, h, 1, dim_v), dtype=torch.float32)
        lse = torch.full((b, h, 1, 1), -torch.finfo(torch.float32).max, device=q.device, dtype=torch.float32)

    if not is_distributed():
        return local_out.to(dtype)

    max_lse = lse.clone()
    dist.all_reduce(max_lse, dist.ReduceOp.MAX)
    den = (lse - max_lse).exp()
    packed = torch.cat((local_out * den, den), dim=-1)  # numerator | denominator in one collective
    dist.all_reduce(packed)
    num, den = packed[..., :-1], packed[..., -1:]
    return (num / den.clamp(min=eps)).to(dtype)
