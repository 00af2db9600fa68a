"""Single-process checks of the portable ops against the dense oracle (runs on CPU)."""
import pytest
import torch

from ring_attention_pytorch_b200 import default_attention, ring_flash_attn
from ring_attention_pytorch_b200.ops.oracle import attention_with_positions
from ring_attention_pytorch_b200.parallel.layout import (from_layout, make_position_map, ring_hop_owners,
                                                         ring_query_owners, to_layout)


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("kmask", [False, True])
@pytest.mark.parametrize("softclamp", [False, True])
def test_naive_flash_matches_dense(causal, kmask, softclamp):
    torch.manual_seed(0)
    q = torch.randn(2, 62, 4, 16, requires_grad=True)
    k = torch.randn(2, 62, 2, 16, requires_grad=True)
    v = torch.randn(2, 62, 2, 16, requires_grad=True)
    mask = (torch.rand(2, 62) > 0.3) if kmask else None
    out = ring_flash_attn(q, k, v, mask, causal, 4, False, False, None, None, softclamp, 30.0)
    ref = default_attention(q, k, v, mask, causal, softclamp, 30.0)
    g = torch.randn_like(out)
    got = torch.autograd.grad(out, (q, k, v), g)
    want = torch.autograd.grad(ref, (q, k, v), g)
    assert torch.allclose(out, ref, atol=2e-6)
    for a, b in zip(got, want):
        assert torch.allclose(a, b, atol=5e-6)


def test_cross_attention_causal_alignment():
    torch.manual_seed(0)
    q, k, v = torch.randn(1, 5, 2, 8), torch.randn(1, 12, 2, 8), torch.randn(1, 12, 2, 8)
    out = ring_flash_attn(q, k, v, None, True, 4)
    ref = default_attention(q, k, v, causal=True)
    assert torch.allclose(out, ref, atol=2e-6)


def test_fully_masked_rows_are_zero_not_nan():
    q, k, v = torch.randn(1, 4, 1, 8), torch.randn(1, 4, 1, 8), torch.randn(1, 4, 1, 8)
    mask = torch.zeros(1, 4, dtype=torch.bool)
    out = ring_flash_attn(q, k, v, mask)
    assert torch.equal(out, torch.zeros_like(out))
    out2, lse = attention_with_positions(q, k, v, key_mask=mask, return_lse=True)
    assert torch.equal(out2, torch.zeros_like(out2)) and torch.isinf(lse).all()


@pytest.mark.parametrize("layout", ["plain", "striped", "zigzag"])
def test_layout_roundtrip_and_positions(layout):
    world, n = 4, 8
    x = torch.arange(world * n).float()[None, :, None]
    y = to_layout(x, layout, world)
    assert torch.equal(from_layout(y, layout, world), x)
    pm = make_position_map(layout, world, n)
    allpos = torch.cat([pm.positions(r) for r in range(world)])
    assert sorted(allpos.tolist()) == list(range(world * n))
    assert torch.equal(y.flatten(), allpos.float())


def test_hop_and_query_owner_schedules():
    pm = make_position_map("plain", 4, 8)
    assert ring_hop_owners(pm, 2, True, None) == [2, 1, 0]
    assert ring_hop_owners(pm, 2, False, None) == [2, 1, 0, 3]
    assert ring_hop_owners(pm, 3, True, 8) == [3, 2]          # look-back of 8 tokens reaches one rank back
    assert ring_query_owners(pm, 1, True, None) == [1, 2, 3]
    pm = make_position_map("striped", 4, 8)
    assert ring_hop_owners(pm, 0, True, None) == [0, 3, 2, 1]
    assert ring_query_owners(pm, 0, True, None) == [0, 1, 2, 3]


def test_lookback_window_matches_oracle():
    torch.manual_seed(1)
    q, k, v = (torch.randn(1, 40, 2, 8) for _ in range(3))
    out = ring_flash_attn(q, k, v, None, True, 8, False, False, 7)
    pos = torch.arange(40)
    ref = attention_with_positions(q, k, v, pos, pos, causal=True, window=7)
    assert torch.allclose(out, ref, atol=2e-6)
