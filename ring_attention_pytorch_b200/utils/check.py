"""Sampled-row verification of a (ring) attention forward + backward at benchmark scale.

A dense fp32 oracle of a 262144-token problem is out of reach for every head, but ONE (batch, head) pair is cheap:
the oracle forward runs in row chunks (lse and O for every query of that head, ~2 S^2 d FLOP), after which

* ``out`` and ``dQ`` of sampled query rows only need those rows' logits, and
* ``dK`` / ``dV`` of sampled key rows only need the matching logit COLUMNS plus lse / delta of every query.

In a ring every rank gathers the chosen head of q, k, v, dO from all ranks (NCCL, cold path), runs the oracle for
the whole head and checks the rows IT owns.  Positions come from the same :class:`PositionMap` the kernels use, so
striped / zig-zag layouts are checked exactly as they are computed.  Used by ``bench.py --check`` and the multi-GPU
tests; the reference has no counterpart (its tests compare small dense problems, assert_flash.py:66-91).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist

from ring_attention_pytorch_b200.parallel.layout import make_position_map


def _gather(t: torch.Tensor, world: int) -> torch.Tensor:
    """[n, ...] on every rank -> [world * n, ...] in rank-major order."""
    if world == 1:
        return t
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t.contiguous())
    return torch.cat(parts, 0)


def _errors(got: torch.Tensor, ref: torch.Tensor) -> Dict[str, float]:
    got, ref = got.float(), ref.float()
    scale = ref.abs().max().clamp_min(1e-20)
    diff = (got - ref).abs()
    # per-element criterion: |err| <= 2 % of the tensor's max magnitude + 6 % of the element's own magnitude
    tol = 0.02 * scale + 0.06 * ref.abs()
    return {
        "max_abs_over_max": float((diff.max() / scale).item()),
        "frac_elements_out_of_tol": float((diff > tol).float().mean().item()),
        "nan": bool(torch.isnan(got).any().item()),
    }


@torch.no_grad()
def sampled_check(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    dout: torch.Tensor,
    out: torch.Tensor,
    dq: torch.Tensor,
    dk: torch.Tensor,
    dv: torch.Tensor,
    *,
    causal: bool,
    layout: str = "plain",
    world: int = 1,
    rank: int = 0,
    batch_index: int = 0,
    head_index: int = 0,
    samples: int = 64,
    chunk: int = 2048,
    seed: int = 0,
    scale: Optional[float] = None,
) -> Dict[str, object]:
    """All tensors are this rank's shards ``[b, n, h(k), d]``.  Returns a dict with per-tensor error statistics
    and ``ok``.  Collective when ``world > 1`` (every rank of the ring must call it)."""
    b, n, h, d = q.shape
    hk = k.shape[2]
    kv_head = head_index % hk  # reference head mapping: query head j uses kv head j % kv_heads
    scale = d ** -0.5 if scale is None else scale
    dev = q.device
    pm = make_position_map(layout, world, n)

    qh = _gather(q[batch_index, :, head_index].float(), world)        # [S, d]
    doh = _gather(dout[batch_index, :, head_index].float(), world)
    kh = _gather(k[batch_index, :, kv_head].float(), world)
    vh = _gather(v[batch_index, :, kv_head].float(), world)
    pos = torch.cat([pm.positions(r, dev) for r in range(world)])     # [S] position of gathered row j
    S = qh.shape[0]

    # oracle forward for the whole head, chunked over query rows
    lse = torch.empty(S, device=dev)
    o_ref = torch.empty(S, d, device=dev)
    for s0 in range(0, S, chunk):
        s1 = min(S, s0 + chunk)
        sim = (qh[s0:s1] @ kh.t()) * scale
        if causal:
            sim.masked_fill_(pos[None, :] > pos[s0:s1, None], float("-inf"))
        l = torch.logsumexp(sim, dim=-1)
        lse[s0:s1] = l
        o_ref[s0:s1] = torch.exp(sim - l[:, None]) @ vh
        del sim
    delta = (o_ref * doh).sum(-1)                                      # [S]

    g = torch.Generator(device="cpu").manual_seed(seed + rank)
    rows = torch.randperm(n, generator=g)[:min(samples, n)].to(dev)   # local row indices checked on this rank
    grow = rows + rank * n                                             # their index in the gathered order

    res: Dict[str, object] = {}
    # ---- out, dQ of the sampled queries -----------------------------------------------------------
    sim = (qh[grow] @ kh.t()) * scale
    if causal:
        sim.masked_fill_(pos[None, :] > pos[grow, None], float("-inf"))
    p = torch.exp(sim - lse[grow, None])
    dp = doh[grow] @ vh.t()
    ds = p * (dp - delta[grow, None])
    dq_ref = (ds @ kh) * scale
    res["out"] = _errors(out[batch_index, rows, head_index], o_ref[grow])
    res["dq"] = _errors(dq[batch_index, rows, head_index], dq_ref)

    # ---- dK, dV of the sampled keys: sum over every query AND every query head of the GQA group ----
    # (only the chosen head's contribution can be formed from one head; with grouped heads the check gathers the
    #  whole group)
    group = [j for j in range(h) if j % hk == kv_head]
    dk_ref = torch.zeros(rows.numel(), d, device=dev)
    dv_ref = torch.zeros(rows.numel(), d, device=dev)
    for j in group:
        if j == head_index:
            qj, doj, lsej, deltaj = qh, doh, lse, delta
        else:
            qj = _gather(q[batch_index, :, j].float(), world)
            doj = _gather(dout[batch_index, :, j].float(), world)
            lsej = torch.empty(S, device=dev)
            oj = torch.empty(S, d, device=dev)
            for s0 in range(0, S, chunk):
                s1 = min(S, s0 + chunk)
                sim = (qj[s0:s1] @ kh.t()) * scale
                if causal:
                    sim.masked_fill_(pos[None, :] > pos[s0:s1, None], float("-inf"))
                l = torch.logsumexp(sim, dim=-1)
                lsej[s0:s1] = l
                oj[s0:s1] = torch.exp(sim - l[:, None]) @ vh
                del sim
            deltaj = (oj * doj).sum(-1)
        simc = (qj @ kh[grow].t()) * scale                             # [S, samples]
        if causal:
            simc.masked_fill_(pos[None, grow] > pos[:, None], float("-inf"))
        pc = torch.exp(simc - lsej[:, None])
        dv_ref += pc.t() @ doj
        dsc = pc * (doj @ vh[grow].t() - deltaj[:, None])
        dk_ref += (dsc.t() @ qj) * scale
    res["dk"] = _errors(dk[batch_index, rows, kv_head], dk_ref)
    res["dv"] = _errors(dv[batch_index, rows, kv_head], dv_ref)

    ok = all((not e["nan"]) and e["max_abs_over_max"] < 3e-2 and e["frac_elements_out_of_tol"] < 1e-3
             for e in res.values())
    if world > 1:
        flag = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item() > 0.5)
    res["ok"] = ok
    res["rows_checked_per_rank"] = int(rows.numel())
    res["head"] = [batch_index, head_index]
    return res
