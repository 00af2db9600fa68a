"""Multi-process (gloo, CPU) tests of the portable ring path and the nn modules.

Oracle = dense fp32 attention over the un-sharded sequence (never the reference's dK/dV, which are wrong –
SURVEY.md §2.8 D1).  Losses are random-cotangent sums, not ``.mean()``, so gradient checks are not vacuous.
"""
import pytest
import torch
import torch.distributed as dist

from dist_utils import run_distributed


# ------------------------------------------------------------------------------------------------
# functional op
# ------------------------------------------------------------------------------------------------
def _op_worker(rank, world, layout, causal, hk, kmask, window):
    from ring_attention_pytorch_b200 import ring_flash_attn
    from ring_attention_pytorch_b200.ops.oracle import attention_with_positions
    from ring_attention_pytorch_b200.parallel.layout import make_position_map

    torch.manual_seed(0)
    b, n, h, d = 2, 12, 4, 8
    qs = [torch.randn(b, n, h, d) for _ in range(world)]
    ks = [torch.randn(b, n, hk, d) for _ in range(world)]
    vs = [torch.randn(b, n, hk, d) for _ in range(world)]
    gs = [torch.randn(b, n, h, d) for _ in range(world)]
    ms = [torch.rand(b, n) > 0.3 for _ in range(world)] if kmask else None
    q, k, v = (t[rank].clone().requires_grad_() for t in (qs, ks, vs))
    out = ring_flash_attn(q, k, v, ms[rank] if kmask else None, causal, 4, True, layout == "striped", window, world,
                          False, 50.0, layout)
    dq, dk, dv = torch.autograd.grad(out, (q, k, v), gs[rank])

    pm = make_position_map(layout, world, n)
    qf = [t.clone().requires_grad_() for t in qs]
    kf = [t.clone().requires_grad_() for t in ks]
    vf = [t.clone().requires_grad_() for t in vs]
    k_all, v_all = torch.cat(kf, 1), torch.cat(vf, 1)
    k_pos = torch.cat([pm.positions(r) for r in range(world)])
    km = torch.cat(ms, 1) if kmask and not causal else None
    loss = 0
    outs = []
    for r in range(world):
        o = attention_with_positions(qf[r], k_all, v_all, pm.positions(r), k_pos, causal=causal, window=window,
                                     key_mask=km)
        outs.append(o)
        loss = loss + (o * gs[r]).sum()
    loss.backward()
    assert torch.allclose(out, outs[rank], atol=2e-5), (out - outs[rank]).abs().max()
    assert torch.allclose(dq, qf[rank].grad, atol=5e-5)
    assert torch.allclose(dk, kf[rank].grad, atol=5e-5), "dK must match the dense oracle (reference defect D1)"
    assert torch.allclose(dv, vf[rank].grad, atol=5e-5), "dV must match the dense oracle (reference defect D1)"


@pytest.mark.parametrize("world,layout,causal,hk,kmask,window", [
    (2, "plain", False, 4, False, None),
    (2, "plain", False, 2, True, None),
    (4, "plain", True, 4, False, None),
    (4, "striped", True, 2, False, None),
    (4, "zigzag", True, 4, False, None),
    (4, "plain", True, 4, False, 10),
    (3, "striped", True, 1, False, 17),
])
def test_ring_flash_attn_distributed(world, layout, causal, hk, kmask, window):
    run_distributed(_op_worker, world, layout, causal, hk, kmask, window)


# ------------------------------------------------------------------------------------------------
# RingAttention / RingTransformer with auto sharding, ring sets, striping, rotary
# ------------------------------------------------------------------------------------------------
def _attn_module_worker(rank, world, causal, striped, num_sharded_batches, rotary, var_batch):
    from math import ceil

    from ring_attention_pytorch_b200 import RingAttention

    torch.manual_seed(0)
    seq_len, dim = 31, 16
    ring_seq_size = ceil(seq_len / world) * num_sharded_batches
    bucket_size = ring_seq_size // 2 if ring_seq_size % 2 == 0 else ring_seq_size
    kw = dict(dim=dim, causal=causal, dim_head=8, heads=4, num_grouped_query_heads=2, bucket_size=bucket_size,
              rotary_embed=rotary, use_cuda_kernel=False)
    ring = RingAttention(ring_attn=True, ring_seq_size=ring_seq_size, striped_ring_attn=striped, auto_shard_seq=True,
                         **kw)
    flash = RingAttention(ring_attn=False, **kw)
    flash.load_state_dict(ring.state_dict())

    batch = 2 + (rank if var_batch else 0)
    if num_sharded_batches > 1:
        batch = 2
    torch.manual_seed(100 + rank)
    x = torch.randn(batch, seq_len, dim)
    g = torch.randn(batch, seq_len, dim)
    xr = x.clone().requires_grad_()
    xf = x.clone().requires_grad_()
    out_r = ring(xr)
    out_f = flash(xf)
    assert torch.allclose(out_r, out_f, atol=2e-5), (out_r - out_f).abs().max()
    (out_r * g).sum().backward()
    (out_f * g).sum().backward()
    # input gradients cross ranks through the auto-shard gather: needs the reduce-scatter backward (D8)
    assert torch.allclose(xr.grad, xf.grad, atol=5e-5), (xr.grad - xf.grad).abs().max()
    # parameter grads: the ring model computes each rank's loss contribution on its shards; summing over
    # ranks must equal the sum of the per-rank dense models
    for (name, pr), (_, pf) in zip(ring.named_parameters(), flash.named_parameters()):
        gr, gf = pr.grad.clone(), pf.grad.clone()
        dist.all_reduce(gr)
        dist.all_reduce(gf)
        assert torch.allclose(gr, gf, atol=2e-4), (name, (gr - gf).abs().max())


@pytest.mark.parametrize("world,causal,striped,nsb,rotary,var_batch", [
    (2, False, False, 1, False, False),
    (2, True, True, 1, True, True),
    (4, True, False, 2, True, False),    # reference fails this (D5)
    (4, True, True, 2, True, False),     # reference fails this (D6)
    (4, True, True, 1, False, True),
])
def test_ring_attention_module(world, causal, striped, nsb, rotary, var_batch):
    run_distributed(_attn_module_worker, world, causal, striped, nsb, rotary, var_batch)


def _transformer_worker(rank, world, causal, striped, num_sharded_batches):
    from math import ceil

    from ring_attention_pytorch_b200 import RingTransformer

    torch.manual_seed(0)
    seq_len = 31
    ring_seq_size = ceil(seq_len / world) * num_sharded_batches
    kw = dict(num_tokens=64, dim=16, depth=2, causal=causal, dim_head=8, heads=4, num_grouped_query_heads=2,
              bucket_size=ring_seq_size, use_cuda_kernel=False)
    ring = RingTransformer(ring_attn=True, striped_ring_attn=striped, ring_seq_size=ring_seq_size, **kw)
    flash = RingTransformer(ring_attn=False, **kw)
    flash.load_state_dict(ring.state_dict())
    torch.manual_seed(100 + rank)
    tokens = torch.randint(0, 64, (2, seq_len))
    g = torch.randn(2, seq_len, 64)
    lr = ring(tokens)
    lf = flash(tokens)
    assert torch.allclose(lr, lf, atol=5e-5), (lr - lf).abs().max()
    (lr * g).sum().backward()
    (lf * g).sum().backward()
    gr, gf = ring.token_emb.weight.grad.clone(), flash.token_emb.weight.grad.clone()
    dist.all_reduce(gr)
    dist.all_reduce(gf)
    assert torch.allclose(gr, gf, atol=5e-4), (gr - gf).abs().max()
    # loss path (labels shifted inside, local mean over the shard)
    loss = ring(tokens, return_loss=True)
    assert torch.isfinite(loss)


@pytest.mark.parametrize("world,causal,striped,nsb", [(2, False, False, 1), (4, True, True, 1), (4, True, True, 2)])
def test_ring_transformer(world, causal, striped, nsb):
    run_distributed(_transformer_worker, world, causal, striped, nsb)


# ------------------------------------------------------------------------------------------------
# tree decode, zig-zag pipeline, all-gather backward
# ------------------------------------------------------------------------------------------------
def _tree_worker(rank, world, seq_len):
    from ring_attention_pytorch_b200 import tree_attn_decode

    torch.manual_seed(0)
    q = torch.randn(1, 8, 1, 16)
    k = torch.randn(1, 8, seq_len, 16)
    v = torch.randn(1, 8, seq_len, 16)
    sim = torch.einsum("bhid,bhjd->bhij", q, k) * 16 ** -0.5
    ref = torch.einsum("bhij,bhjd->bhid", sim.softmax(-1), v)
    out = tree_attn_decode(q, k, v, use_triton=False)
    assert torch.allclose(out, ref, atol=1e-5), (out - ref).abs().max()


@pytest.mark.parametrize("world,seq_len", [(4, 31), (8, 5)])
def test_tree_attn_decode(world, seq_len):
    run_distributed(_tree_worker, world, seq_len)


def _zigzag_worker(rank, world, rotary):
    from ring_attention_pytorch_b200 import RingAttention, apply_rotary_pos_emb, zig_zag_attn, zig_zag_pad_seq, \
        zig_zag_shard

    torch.manual_seed(0)
    seq_len, dim, dim_head, heads = 31, 16, 8, 4
    attn = RingAttention(dim=dim, causal=True, dim_head=dim_head, heads=heads, num_grouped_query_heads=2,
                         ring_attn=False, rotary_embed=rotary, use_cuda_kernel=False)
    torch.manual_seed(7)
    x = torch.randn(2, seq_len, dim)
    g = torch.randn(2, seq_len, dim)
    xl = x.clone().requires_grad_()
    xz = x.clone().requires_grad_()
    ref = attn(xl)

    padded, remove_pad = zig_zag_pad_seq(xz)
    (shard, q_idx, kv_idx), gather_seq = zig_zag_shard(padded, all_gather_batch=False)
    qkv = attn.to_qkv(shard)
    b, n = qkv.shape[:2]
    q, k, v = qkv.view(b, n, -1, dim_head).split(attn.qkv_head_breakdown, dim=-2)
    q, k, v = (t.transpose(1, 2) for t in (q, k, v))
    if rotary:
        pos = attn.rotary_embed(q_idx)
        q = apply_rotary_pos_emb(pos, q, head_dim_first=True)
        k = apply_rotary_pos_emb(pos, k, head_dim_first=True)
    for mode in ("ring", "dense_mask"):
        if mode == "ring":
            o = zig_zag_attn(q, k, v, causal=True)
        else:
            o = zig_zag_attn(q, k, v, attn_mask=q_idx[:, None] >= kv_idx[None, :])
        o = o.transpose(1, 2).reshape(b, n, -1)
        o = attn.to_out(o)
        o = remove_pad(gather_seq(o))
        assert torch.allclose(o, ref, atol=3e-5), (mode, (o - ref).abs().max())
    (o * g).sum().backward()
    (ref * g).sum().backward()
    # x is replicated and every rank back-propagates the same full-sequence loss: the sequence gather's
    # reduce-scatter sums the `world` identical cotangents, and each rank only holds the paths through its
    # own shard, so the true gradient is sum_r(grad_r) / world.
    gz = xz.grad.clone()
    dist.all_reduce(gz)
    gz /= world
    assert torch.allclose(gz, xl.grad, atol=1e-4), (gz - xl.grad).abs().max()


@pytest.mark.parametrize("world,rotary", [(2, False), (4, True)])
def test_zig_zag_pipeline(world, rotary):
    run_distributed(_zigzag_worker, world, rotary)


def _allgather_worker(rank, world):
    from ring_attention_pytorch_b200.parallel.distributed import AllGather

    x = torch.full((rank + 1, 3), float(rank + 1), requires_grad=True)
    y, sizes = AllGather(dim=0)(x)
    assert y.shape[0] == sum(range(1, world + 1)) and sizes.tolist() == list(range(1, world + 1))
    # every rank weights the gathered tensor differently; the correct grad for rank r's rows is the SUM of
    # every rank's weight (reduce-scatter), not just the local one
    (y * float(rank + 1)).sum().backward()
    expect = float(sum(range(1, world + 1)))
    assert torch.allclose(x.grad, torch.full_like(x, expect))


def test_all_gather_backward_is_reduce_scatter():
    run_distributed(_allgather_worker, 3)


# ------------------------------------------------------------------------------------------------
# ring helpers (reference ring.py) with ring sets: world 4 = 2 ring sets x ring size 2
# ------------------------------------------------------------------------------------------------
def _ring_helper_worker(rank, world, ring_size):
    from ring_attention_pytorch_b200.parallel.distributed import all_gather_variable_dim, split_by_rank
    from ring_attention_pytorch_b200.parallel.ring import (all_ring_pass, circular_rank_left, circular_rank_right,
                                                            null_ring_pass, one_ring_pass, ring_pass)

    ring_set, local = divmod(rank, ring_size)
    base = ring_set * ring_size
    assert circular_rank_right(ring_size=ring_size) == base + (local + 1) % ring_size
    assert circular_rank_left(ring_size=ring_size) == base + (local - 1) % ring_size

    x = torch.full((3,), float(rank))
    for num in range(0, ring_size + 1):
        got, sent = ring_pass(num, x.clone(), None, ring_size)
        src = base + (local - num) % ring_size  # data moves `num` positions to the right inside the ring set
        assert torch.equal(got, torch.full((3,), float(src))), (rank, num, got)
        assert torch.equal(sent, x)
    got, _ = one_ring_pass(x.clone(), None, ring_size)
    assert got[0].item() == base + (local - 1) % ring_size

    seen = []
    for info, ((t, none), _) in all_ring_pass(x.clone(), None, ring_size=ring_size):
        assert none is None
        seen.append((info.ring_rank, int(t[0].item()), info.iter_info))
    # the tensor held at step s was produced by the ring-local rank (local - s), of this ring set
    assert [s[0] for s in seen] == [(local - s) % ring_size for s in range(ring_size)]
    assert [s[1] for s in seen] == [base + (local - s) % ring_size for s in range(ring_size)]
    assert seen[0][2] == (True, ring_size == 1) and seen[-1][2] == (ring_size == 1, True)
    limited = list(all_ring_pass(x.clone(), max_iters=1, ring_size=ring_size))
    assert len(limited) == 1 and limited[0][0].iter_info == (True, True)
    assert len(list(null_ring_pass(x))) == 1

    # variable-size gather + split_by_rank round trip
    y = torch.full((rank + 1, 2), float(rank))
    gathered, sizes = all_gather_variable_dim(y, dim=0)
    assert sizes.tolist() == [r + 1 for r in range(world)] and gathered.shape[0] == sum(range(1, world + 1))
    mine, piece_sizes = split_by_rank(gathered.split(sizes.tolist(), dim=0))
    assert torch.equal(mine, y) and piece_sizes.tolist() == sizes.tolist()


def test_ring_helpers_with_ring_sets():
    run_distributed(_ring_helper_worker, 4, 2)
