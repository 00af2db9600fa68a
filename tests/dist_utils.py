"""Spawn ``world_size`` gloo (CPU) or nccl (GPU) processes on 127.0.0.1 and run ``fn(rank, world, *args)``."""
from __future__ import annotations

import os
import socket
import traceback

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _entry(rank, world, port, backend, fn, args, errq):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.set_num_threads(1)
        if backend == "nccl":
            torch.cuda.set_device(rank)
        dist.init_process_group(backend, rank=rank, world_size=world)
        try:
            fn(rank, world, *args)
        finally:
            dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        errq.put((rank, traceback.format_exc()))
        raise


def run_distributed(fn, world: int, *args, backend: str = "gloo", timeout: float = 240.0):
    ctx = mp.get_context("spawn")
    errq = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_entry, args=(r, world, port, backend, fn, args, errq)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout)
    errors = []
    while not errq.empty():
        errors.append(errq.get())
    for p in procs:
        if p.is_alive():
            p.kill()
            errors.append((-1, "timeout"))
    bad = [p.exitcode for p in procs if p.exitcode != 0]
    assert not errors and not bad, "distributed run failed:\n" + "\n".join(f"[rank {r}] {e}" for r, e in errors)
