"""Shape-annotated tensor aliases (reference tensor_typing.py:1-26): ``Float['b n d']`` etc.

Annotation-only sugar built on jaxtyping when it is installed; falls back to plain ``torch.Tensor``.
"""
from __future__ import annotations

from torch import Tensor

try:  # pragma: no cover - optional dependency
    from jaxtyping import Bool as _Bool
    from jaxtyping import Float as _Float
    from jaxtyping import Int as _Int

    class _TorchTyping:
        def __init__(self, abstract_dtype):
            self.abstract_dtype = abstract_dtype

        def __getitem__(self, shapes: str):
            return self.abstract_dtype[Tensor, shapes]

    Float = _TorchTyping(_Float)
    Int = _TorchTyping(_Int)
    Bool = _TorchTyping(_Bool)
except Exception:  # noqa: BLE001

    class _Plain:
        def __getitem__(self, shapes: str):
            return Tensor

    Float = Int = Bool = _Plain()

__all__ = ["Float", "Int", "Bool"]
