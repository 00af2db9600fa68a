"""Property tests (hypothesis) of the host-side ring schedule that both backends and every kernel rely on:
dropping a hop must never drop a visible (query, key) pair, and the K/V-side and Q-side schedules must be transposes
of each other (rank r visits owner o  <=>  owner o is visited by rank r)."""
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from ring_attention_pytorch_b200.parallel.layout import make_position_map, ring_hop_owners, ring_query_owners

cfg = st.tuples(st.sampled_from(["plain", "striped", "zigzag"]), st.integers(1, 8), st.integers(1, 12),
                st.booleans(), st.one_of(st.none(), st.integers(1, 60)))


def _visible(pm, r, o, causal, window):
    q, k = pm.positions(r)[:, None], pm.positions(o)[None, :]
    if not causal:
        return torch.ones(pm.n, pm.n, dtype=torch.bool)
    vis = q >= k
    if window is not None:
        vis &= (q - k) <= window
    return vis


@settings(max_examples=200, deadline=None)
@given(cfg)
def test_dropped_hops_hold_no_visible_pair(c):
    layout, world, half, causal, window = c
    n = 2 * half  # zig-zag needs an even local length
    window = window if causal else None
    pm = make_position_map(layout, world, n)
    for r in range(world):
        hops = ring_hop_owners(pm, r, causal, window)
        assert hops[0] == r and len(set(hops)) == len(hops)
        # ring order: r, r-1, r-2, ... with gaps only where hops were dropped
        steps = [(r - o) % world for o in hops]
        assert steps == sorted(steps)
        for o in range(world):
            if o not in hops:
                assert not _visible(pm, r, o, causal, window).any(), (c, r, o)


@settings(max_examples=200, deadline=None)
@given(cfg)
def test_query_owner_schedule_is_the_transpose(c):
    layout, world, half, causal, window = c
    n = 2 * half
    window = window if causal else None
    pm = make_position_map(layout, world, n)
    visits = {(r, o) for r in range(world) for o in ring_hop_owners(pm, r, causal, window)}
    visited_by = {(r, o) for o in range(world) for r in ring_query_owners(pm, o, causal, window)}
    # the Q-side schedule may be conservative (it can keep a pair the K/V side dropped) but must cover every visit
    assert visits <= visited_by, (c, visits - visited_by)
    for o in range(world):
        assert ring_query_owners(pm, o, causal, window)[0] == o


@settings(max_examples=100, deadline=None)
@given(cfg)
def test_positions_partition_the_sequence(c):
    layout, world, half, _, _ = c
    n = 2 * half
    pm = make_position_map(layout, world, n)
    allpos = torch.cat([pm.positions(r) for r in range(world)])
    assert sorted(allpos.tolist()) == list(range(world * n))
    for r in range(world):
        lo, hi = pm.pos_range(r)
        p = pm.positions(r)
        assert lo == int(p.min()) and hi == int(p.max())
