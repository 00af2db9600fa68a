"""Device-side timing helpers (CUDA events on the launching stream, max over ranks) and NVTX ranges."""
from __future__ import annotations

import contextlib
from typing import Callable, Dict, List

import torch
import torch.distributed as dist


def time_cuda(fn: Callable[[], object], warmup: int = 3, iters: int = 10, flush_l2: bool = False) -> List[float]:
    """Per-iteration milliseconds of ``fn`` measured with CUDA events after ``warmup`` untimed calls."""
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if flush_l2 else None
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()  # 256 MiB > 126 MB L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return out


def max_over_ranks(ms: float) -> float:
    """Reduce a device-timed duration to the slowest rank (multi-GPU numbers are always max over ranks)."""
    if not (dist.is_available() and dist.is_initialized()):
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# Phase timeline: every ``nvtx_range`` of the ring op (pack+barrier / kernel / prep / zero+barrier / convert ...) also
# records a pair of CUDA events while ``enable_phase_timing(True)`` is in effect; ``phase_report()`` turns them into
# per-phase milliseconds of this rank (tools/phase_timeline.py prints the table for every rank of a ring).
_PHASES = None


def enable_phase_timing(on: bool = True) -> None:
    global _PHASES
    _PHASES = [] if on else None


def phase_report(reset: bool = True) -> Dict[str, List[float]]:
    """{phase name: [ms of every occurrence]} since phase timing was enabled (synchronises the device)."""
    global _PHASES
    out: Dict[str, List[float]] = {}
    if _PHASES is None:
        return out
    torch.cuda.synchronize()
    for name, e0, e1 in _PHASES:
        out.setdefault(name, []).append(e0.elapsed_time(e1))
    if reset:
        _PHASES = []
    return out


@contextlib.contextmanager
def nvtx_range(name: str):
    on_gpu = torch.cuda.is_available()
    if on_gpu:
        torch.cuda.nvtx.range_push(name)
    ev = None
    if _PHASES is not None and on_gpu:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    try:
        yield
    finally:
        if ev is not None:
            ev[1].record()
            _PHASES.append((name, ev[0], ev[1]))
        if on_gpu:
            torch.cuda.nvtx.range_pop()


def attention_flops(batch: int, heads: int, seq_q: int, seq_k: int, dim: int, causal: bool, backward: bool = False):
    """Algorithmic FLOPs: fwd = 4*b*h*nq*nk*d (x0.5 causal); bwd = 2.5x fwd (5 GEMMs)."""
    f = 4.0 * batch * heads * seq_q * seq_k * dim * (0.5 if causal else 1.0)
    return f * 2.5 if backward else f
