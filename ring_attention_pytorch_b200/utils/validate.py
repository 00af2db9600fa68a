"""Runtime argument validation for the public constructors and ops.

The reference type-checks every public entry point with ``beartype`` (ring_attention.py:47, 103, 284, 489;
ring_flash_attention.py:391).  ``typecheck`` is that decorator when beartype is importable (it is in the B200 image and
listed in ``requirements.txt``) and a no-op otherwise, so the package stays importable on a bare PyTorch install.  On top
of the annotation checks, ``check_attention_inputs`` validates what annotations cannot express: tensor ranks, matching
batch / head-dim sizes, grouped-query divisibility and mask shapes — failing with a message that names the argument
instead of a kernel-launch assertion.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

try:  # pragma: no cover - exercised implicitly by every decorated call
    from beartype import BeartypeConf
    from beartype import beartype as _beartype

    # PEP 484 numeric tower: an int is accepted where a float is annotated (``theta=10000``, ``softclamp_value=50``)
    _checked = _beartype(conf=BeartypeConf(is_pep484_tower=True))

    def typecheck(fn):
        return _checked(fn)

    HAVE_BEARTYPE = True
except Exception:  # noqa: BLE001

    def typecheck(fn):
        return fn

    HAVE_BEARTYPE = False


def check_attention_inputs(q: Tensor, k: Tensor, v: Tensor, mask: Optional[Tensor] = None, *, name: str = "attention",
                           head_dim_first: bool = False, max_head_dim: Optional[int] = None) -> None:
    """q [b, n, h, d], k / v [b, n_k, hk, d] (or head-first).  Raises ValueError naming the offending argument."""
    for nm, t in (("q", q), ("k", k), ("v", v)):
        if not torch.is_tensor(t) or t.dim() != 4:
            raise ValueError(f"{name}: {nm} must be a 4-D tensor, got {tuple(t.shape) if torch.is_tensor(t) else type(t)}")
    hd, sd = (1, 2) if head_dim_first else (2, 1)
    if k.shape != v.shape:
        raise ValueError(f"{name}: k and v must have the same shape, got {tuple(k.shape)} and {tuple(v.shape)}")
    if q.shape[0] != k.shape[0]:
        raise ValueError(f"{name}: batch sizes differ: q {q.shape[0]}, k {k.shape[0]}")
    if q.shape[3] != k.shape[3]:
        raise ValueError(f"{name}: head dims differ: q {q.shape[3]}, k {k.shape[3]}")
    if q.shape[hd] % k.shape[hd] != 0:
        raise ValueError(f"{name}: query heads ({q.shape[hd]}) must be a multiple of key/value heads ({k.shape[hd]})")
    if max_head_dim is not None and q.shape[3] > max_head_dim:
        raise ValueError(f"{name}: head dim {q.shape[3]} exceeds the supported maximum {max_head_dim}")
    if not (q.device == k.device == v.device):
        raise ValueError(f"{name}: q, k, v must live on one device, got {q.device}, {k.device}, {v.device}")
    if mask is not None:
        if mask.dtype != torch.bool or mask.dim() != 2 or mask.shape[0] != k.shape[0] or mask.shape[1] != k.shape[sd]:
            raise ValueError(f"{name}: mask must be bool [batch, keys] = [{k.shape[0]}, {k.shape[sd]}], got "
                             f"{mask.dtype} {tuple(mask.shape)}")
