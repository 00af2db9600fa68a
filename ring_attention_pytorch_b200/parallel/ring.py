"""Ring topology arithmetic and the host-driven P2P ring used by the pure-PyTorch (CPU / gloo) path.

The sm_100a path does not use this module on its hot path – there the "ring" is a schedule evaluated
inside the kernel and K/V move with in-kernel bulk-TMA copies over NVLink (csrc/attn_fwd_sm100.cu).
This module keeps the reference's functional surface (reference ring.py:27-124) for the portable path
and for users that built on it, with these differences:

* no ``dist.barrier()`` after every exchange (reference ring.py:57-60) – ``batch_isend_irecv`` + wait is
  already a complete point-to-point synchronisation;
* ``ring_pass`` honours ``num_ring_passes`` (reference ring.py:62-77 ignores it);
* ring sets (``ring_size < world_size``) always use ring-local arithmetic.
"""
from __future__ import annotations

from collections import namedtuple
from typing import Optional

import torch
import torch.distributed as dist
from torch import Tensor

from ring_attention_pytorch_b200.parallel.distributed import default, exists, get_rank, get_world_size


def circular_index_left(pos: int, ring_size: int, num: int = 1) -> int:
    return ((pos - num) + ring_size) % ring_size


def circular_index_right(pos: int, ring_size: int, num: int = 1) -> int:
    return (pos + num) % ring_size


def circular_rank_left(rank: Optional[int] = None, ring_size: Optional[int] = None, num: int = 1) -> int:
    rank = default(rank, get_rank())
    ring_size = default(ring_size, get_world_size())
    ring_set_num = rank // ring_size
    offset = ring_set_num * ring_size
    return circular_index_left(rank, ring_size, num) + offset


def circular_rank_right(rank: Optional[int] = None, ring_size: Optional[int] = None, num: int = 1) -> int:
    rank = default(rank, get_rank())
    ring_size = default(ring_size, get_world_size())
    ring_set_num = rank // ring_size
    offset = ring_set_num * ring_size
    return circular_index_right(rank, ring_size, num) + offset


def send_and_receive_(x: Tensor, receive_buffer: Tensor, send_to_rank: int, receive_from_rank: int) -> None:
    """One ring exchange (reference ring.py:51-60, minus the global barrier)."""
    ops = [dist.P2POp(dist.isend, x, send_to_rank), dist.P2POp(dist.irecv, receive_buffer, receive_from_rank)]
    for req in dist.batch_isend_irecv(ops):
        req.wait()


def ring_pass(num_ring_passes: int, x: Tensor, receive_buffer: Optional[Tensor] = None, ring_size: Optional[int] = None):
    """Move ``x`` ``num_ring_passes`` positions to the right around this rank's ring set.

    Returns ``(received, sent)`` like the reference so the sent tensor can be reused as the next receive
    buffer (reference ring.py:62-77).
    """
    ring_size = default(ring_size, get_world_size())
    x = x.contiguous()
    if not exists(receive_buffer):
        receive_buffer = torch.zeros_like(x)
    else:
        receive_buffer = receive_buffer.contiguous()
    num = num_ring_passes % ring_size if ring_size > 0 else 0
    if num == 0:
        receive_buffer.copy_(x)
        return receive_buffer, x
    left = circular_rank_left(ring_size=ring_size, num=num)
    right = circular_rank_right(ring_size=ring_size, num=num)
    send_and_receive_(x, receive_buffer, right, left)
    return receive_buffer, x


def one_ring_pass(x: Tensor, receive_buffer: Optional[Tensor] = None, ring_size: Optional[int] = None):
    return ring_pass(1, x, receive_buffer, ring_size)


RingInfo = namedtuple("RingInfo", ["ring_rank", "iter_info"])


def null_ring_pass(*tensors, max_iters=None, receive_buffers=None, ring_size=None):
    """reference ring.py:85-86"""
    yield RingInfo(0, (True, True)), (tensors, receive_buffers)


def all_ring_pass(*tensors, max_iters: Optional[int] = None, receive_buffers=None, ring_size: Optional[int] = None):
    """Iterate over the ring: yields the tensors currently held together with the ring-local rank of the
    rank that produced them (reference ring.py:88-124; here ``ring_rank`` is always ring-local)."""
    ring_size = default(ring_size, get_world_size())
    max_iters = default(max_iters, ring_size)
    receive_buffers = default(receive_buffers, (None,) * len(tensors))
    total_iters = max(1, min(ring_size, max_iters))

    curr_ring_pos = get_rank() % ring_size
    for ind in range(total_iters):
        is_first, is_last = ind == 0, ind == total_iters - 1
        yield RingInfo(curr_ring_pos, (is_first, is_last)), (tensors, receive_buffers)
        curr_ring_pos = circular_index_left(curr_ring_pos, ring_size)
        if is_last:
            continue
        new_tensors, new_buffers = [], []
        for tensor, buffer in zip(tensors, receive_buffers):
            if not exists(tensor):
                new_tensors.append(None)
                new_buffers.append(None)
                continue
            new_tensor, new_buffer = one_ring_pass(tensor, buffer, ring_size)
            new_tensors.append(new_tensor)
            new_buffers.append(new_buffer)
        tensors, receive_buffers = new_tensors, new_buffers
