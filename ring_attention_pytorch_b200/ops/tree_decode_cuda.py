"""sm_100a tree-attention decode (see csrc/tree_decode_sm100.cu and ops/tree_decode.py)."""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from ring_attention_pytorch_b200.ops import _ext
from ring_attention_pytorch_b200.parallel.distributed import get_world_size, is_distributed

LAUNCHES = {"count": 0}


def _choose_splits(n: int, ctas_per_split: int, sm_count: int = 148) -> int:
    if n <= 0:
        return 1
    want = max(1, (2 * sm_count + ctas_per_split - 1) // ctas_per_split)
    return max(1, min(want, (n + 255) // 256))


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
) -> Tensor:
    """q [b, h, 1, d]; k, v [b, hk, n, d] this rank's shard (bf16 / fp16 / float8_e4m3fn) or None.

    ``k_scale`` / ``v_scale``: optional fp32 dequantisation scales for the fp8 path, either per (batch, kv head)
    (``numel == b*hk``) or block-scaled ``[b*hk, n_blocks]`` with one scale per ``scale_block_keys`` keys
    (a multiple of 64).
    Returns [b, h, 1, d] in q's dtype (fp32 if q is fp32).
    """
    ops = _ext.ops()
    b, h, _, d = q.shape
    assert dim_v == d, "the decode kernel assumes dim_v == dim_qk"
    dev = q.device
    qf = q.reshape(b, h, d).float().contiguous()
    n = 0
    hk = h
    if k is not None and k.shape[-2] > 0:
        if k.dtype == torch.float32:
            k, v = k.to(torch.bfloat16), v.to(torch.bfloat16)
        k, v = k.contiguous(), v.contiguous()
        hk, n = k.shape[1], k.shape[2]
    else:
        k = v = None
    g = h // hk
    zchunks = (g + 3) // 4
    splits = _choose_splits(n, b * hk * zchunks)
    scratch = torch.empty(b * hk * splits * g * (d + 2), dtype=torch.float32, device=dev)
    out_dtype = q.dtype if q.dtype in (torch.bfloat16, torch.float16) else torch.float32
    out = torch.empty(b, h, d, dtype=out_dtype, device=dev)
    nbytes = b * h * (d + 2) * 4
    scale = d ** -0.5
    if is_distributed():
        from ring_attention_pytorch_b200.parallel.symm import get_workspace

        ws = get_workspace(get_world_size(), dev)
        stage, peer_ptrs = ws.staging("tree_partial", nbytes)
        partial = stage.view(torch.float32)
        ops.tree_decode_partial(qf, k, v, k_scale, v_scale, scratch, partial, hk, splits, scale, scale_block_keys)
        ws.barrier()
        ops.tree_decode_reduce(peer_ptrs, out, eps)
        LAUNCHES["count"] += 4 if n > 0 else 3
    else:
        partial = torch.empty(b * h * (d + 2), dtype=torch.float32, device=dev)
        ops.tree_decode_partial(qf, k, v, k_scale, v_scale, scratch, partial, hk, splits, scale, scale_block_keys)
        ops.tree_decode_reduce([partial.data_ptr()], out, eps)
        LAUNCHES["count"] += 3 if n > 0 else 2
    return out.view(b, h, 1, d)
