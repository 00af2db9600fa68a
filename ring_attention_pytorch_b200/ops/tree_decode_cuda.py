"""sm_100a tree-attention decode: ONE persistent cooperative kernel per rank and step (``csrc/tree_decode_sm100.cu``).

The kernel computes the split-KV partials of this rank's shard, merges the splits, publishes ``(out, lse)`` in a
symmetric buffer, signals the peers and merges all ranks' partials — over NVLink peer loads, or, when the buffers have a
multicast (NVLS) mapping, with ``multimem.ld_reduce`` inside the NVSwitch.  This wrapper only owns the buffers: they are
cached per (device, world, rows, head dim), every counter is self-resetting and the cross-rank epoch lives in device
memory, so a decode step allocates nothing and is CUDA-graph capturable.  Reference: tree_attn_decoding.py:60-102 (one
Triton launch padded to 128 rows + three all-reduces).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

from ring_attention_pytorch_b200.ops import _ext
from ring_attention_pytorch_b200.parallel.distributed import get_rank, get_world_size, is_distributed

LAUNCHES = {"count": 0}
# nvls "auto": use the NVSwitch multicast mapping when torch's symmetric memory can provide one, else NVLink peer loads
# tensor_core "auto": head dim 128 shards of at least one 128-key tile run the tcgen05 kernel (K / V tiles go from TMA
#                     straight into the MMA; fp8 caches use kind::f8f6f4), everything else the CUDA-core kernel
CONFIG = {"nvls": "auto", "tensor_core": "auto"}
K_MAX_WORLD = 16
PAD_WORDS = 2 * K_MAX_WORLD  # two signal rounds


def _choose_splits(n: int, groups: int, resident_ctas: int) -> int:
    """Enough work units to fill the persistent grid about twice, at least 256 keys per split."""
    if n <= 0:
        return 1
    want = max(1, (2 * resident_ctas + groups - 1) // groups)
    return max(1, min(want, (n + 255) // 256))


@dataclass
class _Buffers:
    rows: int
    d: int
    world: int
    rank: int
    partial_ptrs: List[int]
    aux_local_ptr: int
    pad_ptrs: List[int]
    mc_partial_ptr: int
    mc_aux_ptr: int
    counters: Tensor
    keep: tuple  # owners of the memory above
    scratch: Optional[Tensor] = None
    group_done: Optional[Tensor] = None


_cache: Dict[Tuple, _Buffers] = {}


def _layout(rows: int, d: int) -> Tuple[int, int, int, int]:
    """Byte offsets of (partials, aux, pads) inside one symmetric allocation and its total size."""
    partial_bytes = 2 * rows * (d + 4) * 4
    aux_bytes = 2 * 2 * rows * 4
    pad_bytes = PAD_WORDS * 4
    a = (partial_bytes + 255) // 256 * 256
    b = a + (aux_bytes + 255) // 256 * 256
    return 0, a, b, b + (pad_bytes + 255) // 256 * 256


def _alloc_symmetric(nbytes: int, dev: torch.device, world: int, rank: int):
    """(local uint8 tensor, per-rank base addresses, multicast base or 0, owner).  Tries torch's symmetric memory first
    (it can bind the allocation to an NVSwitch multicast object); falls back to the package's own IPC regions."""
    if world == 1:
        t = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        return t, [t.data_ptr()], 0, t
    if CONFIG["nvls"] in ("auto", True, "on"):
        try:
            import torch.distributed._symmetric_memory as symm_mem

            t = symm_mem.empty(nbytes, dtype=torch.uint8, device=dev)
            hdl = symm_mem.rendezvous(t, dist.group.WORLD)
            t.zero_()
            torch.cuda.synchronize(dev)
            dist.barrier()
            mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
            return t, [int(x) for x in hdl.buffer_ptrs], mc, (t, hdl)
        except Exception as e:  # noqa: BLE001 - no fabric / multicast support: plain peer mappings still work
            if CONFIG["nvls"] in (True, "on"):
                raise
            _alloc_symmetric.last_error = f"{type(e).__name__}: {e}"
    from ring_attention_pytorch_b200.parallel.symm import get_workspace

    ws = get_workspace(world, dev)
    reg = ws.region(f"tree_decode_{nbytes}", nbytes)
    reg.local.zero_()
    torch.cuda.synchronize(dev)
    dist.barrier()
    return reg.local, list(reg.peer_ptrs), 0, reg


_alloc_symmetric.last_error = None


def _buffers(rows: int, d: int, dev: torch.device) -> _Buffers:
    world = get_world_size() if is_distributed() else 1
    rank = get_rank() if world > 1 else 0
    key = (dev.index, world, rows, d)
    buf = _cache.get(key)
    if buf is None:
        off_p, off_a, off_s, total = _layout(rows, d)
        local, bases, mc, owner = _alloc_symmetric(total, dev, world, rank)
        buf = _Buffers(rows=rows, d=d, world=world, rank=rank, partial_ptrs=[x + off_p for x in bases],
                       aux_local_ptr=bases[rank] + off_a, pad_ptrs=[x + off_s for x in bases],
                       mc_partial_ptr=(mc + off_p) if mc else 0, mc_aux_ptr=(mc + off_a) if mc else 0,
                       counters=torch.zeros(4, dtype=torch.int32, device=dev), keep=(local, owner))
        _cache[key] = buf
    return buf


def uses_nvls(q: Tensor) -> bool:
    """True when the decode of this (shape, device) merges through the NVSwitch multicast mapping."""
    b, h, _, d = q.shape
    key = (q.device.index, get_world_size() if is_distributed() else 1, b * h, d)
    return key in _cache and _cache[key].mc_partial_ptr != 0


def _is_cache_prefix(t: Tensor) -> bool:
    """[b, hk, n, d] with dense rows and uniformly strided (batch, head) planes: the filled prefix of a growing
    [b, hk, capacity, d] cache.  The tensor-core kernel reads it in place (tensor-map plane stride), no copy."""
    b, hk, n, d = t.shape
    sb, sh, sn, sd = t.stride()
    if t.is_contiguous():
        return True
    return (b > 1 and hk > 1 and sd == 1 and sn == d and sb == hk * sh and sh >= n * d
            and (sh * t.element_size()) % 16 == 0)


@torch.no_grad()
def tree_decode_cuda(
    q: Tensor,
    k: Optional[Tensor],
    v: Optional[Tensor],
    *,
    dim_v: int,
    eps: float = 1e-8,
    k_scale: Optional[Tensor] = None,
    v_scale: Optional[Tensor] = None,
    scale_block_keys: int = 0,
    out: Optional[Tensor] = None,
) -> Tensor:
    """q [b, h, 1, d] (bf16 / fp16 / fp32); k, v [b, hk, n, d] this rank's shard (bf16 / fp16 / float8_e4m3fn) or None.
    k / v may be the filled prefix ``cache[:, :, :n]`` of a larger ``[b, hk, capacity, d]`` buffer: the tensor-core kernel
    (head dim 128, n >= 128) reads it in place; other cases are made contiguous first.

    ``k_scale`` / ``v_scale``: optional fp32 dequantisation scales for the fp8 path, either per (batch, kv head)
    (``numel == b*hk``) or block-scaled ``[b*hk, n_blocks]`` with one scale per ``scale_block_keys`` keys
    (a multiple of 64).  ``out`` ([b, h, 1, d]) may be passed to make the call allocation free (CUDA graphs).
    Returns [b, h, 1, d] in q's dtype.
    """
    ops = _ext.ops()
    b, h, _, d = q.shape
    assert dim_v == d, "the decode kernel assumes dim_v == dim_qk"
    dev = q.device
    q3 = q.reshape(b, h, d)
    if not q3.is_contiguous():
        q3 = q3.contiguous()
    if q3.dtype not in (torch.bfloat16, torch.float16, torch.float32):
        q3 = q3.float()
    n, hk = 0, h
    if k is not None and k.shape[-2] > 0:
        if k.dtype == torch.float32:
            k, v = k.to(torch.bfloat16), v.to(torch.bfloat16)
        hk, n = k.shape[1], k.shape[2]
    else:
        k = v = None
    g = h // hk
    kv_kind = 0 if k is None or k.dtype == torch.bfloat16 else (1 if k.dtype == torch.float16 else 2)
    tc = CONFIG["tensor_core"]
    # a 128-key tile of the tensor-core kernel must lie inside one scale block
    use_tc = tc in ("auto", True, "on") and d == 128 and n >= 128 and scale_block_keys % 128 == 0
    if k is not None and not (use_tc and _is_cache_prefix(k) and v.stride() == k.stride()):
        k, v = k.contiguous(), v.contiguous()  # no-op for dense inputs
    gm = (4 if g <= 4 else 16) if use_tc else 4  # query heads per work unit (tensor-core kernel: template bound GM)
    groups = b * hk * ((g + gm - 1) // gm)
    resident = int(ops.tree_decode_max_ctas(d, kv_kind, use_tc))
    splits = _choose_splits(n, groups, resident)
    buf = _buffers(b * h, d, dev)
    need = b * hk * splits * g * (d + 4)
    if buf.scratch is None or buf.scratch.numel() < need:
        buf.scratch = torch.empty(need, dtype=torch.float32, device=dev)
    if buf.group_done is None or buf.group_done.numel() < b * hk * ((g + 3) // 4):
        buf.group_done = torch.zeros(b * hk * ((g + 3) // 4), dtype=torch.int32, device=dev)
    if out is None:
        out_dtype = q.dtype if q.dtype in (torch.bfloat16, torch.float16) else torch.float32
        out = torch.empty(b, h, 1, d, dtype=out_dtype, device=dev)
    units = groups * splits if n > 0 else 0
    grid = max(1, min(resident, max(units, (b * h + 3) // 4)))
    ops.tree_decode(q3, k, v, k_scale, v_scale, buf.scratch, buf.group_done, buf.counters, buf.partial_ptrs,
                    buf.aux_local_ptr, buf.pad_ptrs, buf.mc_partial_ptr, buf.mc_aux_ptr, buf.rank, out.view(b, h, d), hk,
                    splits, d ** -0.5, scale_block_keys, eps, grid, use_tc)
    LAUNCHES["count"] += 1
    return out
