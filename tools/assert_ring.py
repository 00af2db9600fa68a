#!/usr/bin/env python
"""Command-line parity checks, the counterpart of the reference's ``assert.py`` / ``assert_attn.py`` /
``assert_flash.py`` / ``assert_tree_attn.py`` / ``assert_zig_zag.py`` drivers (same flag names where they exist).

    python tools/assert_ring.py transformer --world-size 4 --causal --striped-ring-attn --seq-len 31
    python tools/assert_ring.py attn --world-size 4 --num-sharded-batches 2 --causal
    python tools/assert_ring.py flash --causal --rand-key-pad-mask --softclamp-qk-sim
    python tools/assert_ring.py tree --world-size 8 --seq-len 5
    python tools/assert_ring.py zigzag --world-size 4 --seq-len 61
    python tools/assert_ring.py transformer --world-size 2 --use-cuda --causal --striped-ring-attn   # on a GPU box

Unlike the reference drivers the oracle is the dense fp32 model, gradients are compared with random-cotangent
sum losses (not ``.mean()``), and tolerances scale with the compute dtype.
"""
from __future__ import annotations

import os
import socket
import sys
from math import ceil

import click
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _setup(rank, world, port, use_cuda):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if use_cuda:
        torch.cuda.set_device(rank)
    dist.init_process_group("nccl" if use_cuda else "gloo", rank=rank, world_size=world)


def _check(name, a, b, atol):
    err = (a.float() - b.float()).abs().max().item()
    ok = err <= atol
    if dist.get_rank() == 0:
        print(("✅" if ok else "❌") + f" {name}: max abs err {err:.3e} (atol {atol:g})")
    assert ok, name


def _model_worker(rank, world, port, kind, o):
    from ring_attention_pytorch_b200 import RingAttention, RingTransformer

    use_cuda = o["use_cuda"]
    _setup(rank, world, port, use_cuda)
    dev = torch.device("cuda", rank) if use_cuda else torch.device("cpu")
    torch.manual_seed(0)
    ring_seq_size = ceil(o["seq_len"] / world) * o["num_sharded_batches"]
    bucket = max(1, ring_seq_size // o["num_buckets"])
    if ring_seq_size % bucket:
        bucket = ring_seq_size
    common = dict(dim=o["model_dim"], causal=o["causal"], dim_head=o["dim_head"], heads=o["heads"],
                  num_grouped_query_heads=o["num_grouped_query_heads"], bucket_size=bucket)
    if kind == "transformer":
        ring = RingTransformer(num_tokens=256, depth=2, ring_attn=True, striped_ring_attn=o["striped_ring_attn"],
                               ring_seq_size=ring_seq_size, use_cuda_kernel=use_cuda, **common)
        twin = RingTransformer(num_tokens=256, depth=2, ring_attn=False, use_cuda_kernel=False,
                               force_regular_attn=o["compare_regular_attn"] or use_cuda, **common)
    else:
        ring = RingAttention(ring_attn=True, striped_ring_attn=o["striped_ring_attn"], ring_seq_size=ring_seq_size,
                             auto_shard_seq=True, rotary_embed=True, use_cuda_kernel=use_cuda, **common)
        twin = RingAttention(ring_attn=False, rotary_embed=True, use_cuda_kernel=False,
                             force_regular_attn=o["compare_regular_attn"] or use_cuda, **common)
    twin.load_state_dict(ring.state_dict())
    ring, twin = ring.to(dev), twin.to(dev)
    batch = o["batch_size"] + (rank if o["batch_size_var_len"] else 0)
    torch.manual_seed(100 + rank)
    if kind == "transformer":
        x = torch.randint(0, 256, (batch, o["seq_len"]), device=dev)
        xr = xt = x
    else:
        x = torch.randn(batch, o["seq_len"], o["model_dim"], device=dev)
        xr, xt = x.clone().requires_grad_(), x.clone().requires_grad_()
    ctx = torch.autocast("cuda", dtype=torch.bfloat16) if use_cuda else torch.autocast("cpu", enabled=False)
    with ctx:
        out_r, out_t = ring(xr), twin(xt)
    g = torch.randn_like(out_t.float())
    (out_r.float() * g).sum().backward()
    (out_t.float() * g).sum().backward()
    tol_out, tol_grad = (8e-2, 8e-2) if use_cuda else (5e-5, 5e-4)
    _check("output", out_r, out_t, tol_out)
    if kind == "transformer":
        gr, gt = ring.token_emb.weight.grad.clone(), twin.token_emb.weight.grad.clone()
        dist.all_reduce(gr)
        dist.all_reduce(gt)
        _check("token_emb grad (summed over ranks)", gr, gt, tol_grad * max(1.0, gt.abs().max().item()))
    else:
        _check("input grad", xr.grad, xt.grad, tol_grad * max(1.0, xt.grad.abs().max().item()))
    dist.destroy_process_group()


def _tree_worker(rank, world, port, o):
    from ring_attention_pytorch_b200 import tree_attn_decode

    _setup(rank, world, port, o["use_cuda"])
    dev = torch.device("cuda", rank) if o["use_cuda"] else torch.device("cpu")
    dt = torch.bfloat16 if o["use_cuda"] else torch.float32
    torch.manual_seed(0)
    q = torch.randn(o["batch_size"], o["heads"], 1, o["dim_head"], device=dev, dtype=dt)
    k = torch.randn(o["batch_size"], o["heads"], o["seq_len"], o["dim_head"], device=dev, dtype=dt)
    v = torch.randn(o["batch_size"], o["heads"], o["seq_len"], o["dim_head"], device=dev, dtype=dt)
    sim = torch.einsum("bhid,bhjd->bhij", q.float(), k.float()) * o["dim_head"] ** -0.5
    ref = torch.einsum("bhij,bhjd->bhid", sim.softmax(-1), v.float())
    out = tree_attn_decode(q, k, v)
    _check("tree attention decode", out, ref, 2e-2 if o["use_cuda"] else 1e-5)
    dist.destroy_process_group()


def _zigzag_worker(rank, world, port, o):
    """Counterpart of the reference's assert_zig_zag.py:99-131: a causal attention layer evaluated on zig-zag shards
    (pad -> shard -> attn -> inverse) against the same layer on the full sequence, outputs and input gradients."""
    from ring_attention_pytorch_b200 import default_attention, zig_zag_attn, zig_zag_pad_seq, zig_zag_shard

    _setup(rank, world, port, o["use_cuda"])
    dev = torch.device("cuda", rank) if o["use_cuda"] else torch.device("cpu")
    dt = torch.bfloat16 if o["use_cuda"] else torch.float32
    torch.manual_seed(0)
    b, h, hk, d, n = o["batch_size"], o["heads"], max(1, o["heads"] // o["num_grouped_query_heads"]), o["dim_head"], o["seq_len"]
    q = torch.randn(b, h, n, d, device=dev, dtype=dt)
    k = torch.randn(b, hk, n, d, device=dev, dtype=dt)
    v = torch.randn(b, hk, n, d, device=dev, dtype=dt)
    g = torch.randn(b, h, n, d, device=dev, dtype=dt)
    qf, kf, vf = (t.detach().float().clone().requires_grad_() for t in (q, k, v))
    ref = default_attention(qf.transpose(1, 2), kf.transpose(1, 2), vf.transpose(1, 2), causal=True).transpose(1, 2)
    (ref * g.float()).sum().backward()

    qz, kz, vz = (t.detach().clone().requires_grad_() for t in (q, k, v))
    padded = [zig_zag_pad_seq(t)[0] for t in (qz, kz, vz)]
    (ql, _, _), inverse = zig_zag_shard(padded[0])
    (kl, _, _), _ = zig_zag_shard(padded[1])
    (vl, _, _), _ = zig_zag_shard(padded[2])
    out = inverse(zig_zag_attn(ql, kl, vl, causal=True))[..., :n, :]
    (out.float() * g.float()).sum().backward()
    tol = 5e-2 if o["use_cuda"] else 1e-4
    _check("zig-zag output", out, ref, tol)
    # Every rank evaluates the same replicated loss on the re-assembled output, so the all-gather's backward hands each
    # shard the sum over ranks (world x the true gradient), and a rank only touches its own two chunks of its input
    # copy: the full gradient is the sum over ranks divided by the world size.
    for t in (qz, kz, vz):
        dist.all_reduce(t.grad)
        t.grad /= world
    _check("zig-zag dq", qz.grad, qf.grad, tol * max(1.0, qf.grad.abs().max().item()))
    _check("zig-zag dk", kz.grad, kf.grad, tol * max(1.0, kf.grad.abs().max().item()))
    _check("zig-zag dv", vz.grad, vf.grad, tol * max(1.0, vf.grad.abs().max().item()))
    dist.destroy_process_group()


@click.command()
@click.argument("what", type=click.Choice(["transformer", "attn", "flash", "tree", "zigzag"]))
@click.option("--world-size", default=2)
@click.option("--batch-size", default=2)
@click.option("--num-sharded-batches", default=1)
@click.option("--batch-size-var-len", is_flag=True)
@click.option("--use-cuda", is_flag=True)
@click.option("--causal", is_flag=True)
@click.option("--striped-ring-attn", is_flag=True)
@click.option("--num-buckets", default=2)
@click.option("--seq-len", default=31)
@click.option("--model-dim", default=16)
@click.option("--heads", default=4)
@click.option("--num-grouped-query-heads", default=2)
@click.option("--dim-head", default=8)
@click.option("--compare-regular-attn", is_flag=True)
@click.option("--rand-key-pad-mask", is_flag=True)
@click.option("--softclamp-qk-sim", is_flag=True)
def main(what, **o):
    if o["use_cuda"]:
        assert torch.cuda.device_count() >= o["world_size"], "not enough GPUs"
        if o["dim_head"] < 64:
            o["dim_head"], o["model_dim"] = 64, max(o["model_dim"], 128)
    if what == "flash":
        from ring_attention_pytorch_b200 import default_attention, ring_flash_attn

        torch.manual_seed(0)
        q = torch.randn(2, o["seq_len"], o["heads"], o["dim_head"], requires_grad=True)
        k = torch.randn(2, o["seq_len"], o["heads"] // o["num_grouped_query_heads"], o["dim_head"], requires_grad=True)
        v = torch.randn_like(k, requires_grad=True)
        mask = (torch.rand(2, o["seq_len"]) > 0.3) if o["rand_key_pad_mask"] else None
        a = ring_flash_attn(q, k, v, mask, o["causal"], 4, False, False, None, None, o["softclamp_qk_sim"], 50.0)
        b = default_attention(q, k, v, mask, o["causal"], o["softclamp_qk_sim"], 50.0)
        g = torch.randn_like(a)
        ga, gb = torch.autograd.grad(a, (q, k, v), g), torch.autograd.grad(b, (q, k, v), g)
        assert torch.allclose(a, b, atol=2e-6)
        assert all(torch.allclose(x, y, atol=5e-6) for x, y in zip(ga, gb))
        print("✅ flash attention output and dq/dk/dv match the dense oracle")
        return
    port = _free_port()
    if what == "tree":
        mp.spawn(_tree_worker, args=(o["world_size"], port, o), nprocs=o["world_size"], join=True)
    elif what == "zigzag":
        mp.spawn(_zigzag_worker, args=(o["world_size"], port, o), nprocs=o["world_size"], join=True)
    else:
        mp.spawn(_model_worker, args=(o["world_size"], port, what, o), nprocs=o["world_size"], join=True)


if __name__ == "__main__":
    main()
