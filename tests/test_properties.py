"""Property tests (hypothesis) for the host-side pieces and the oracle's invariants (CPU)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from nerrf_b200 import graph as G
from nerrf_b200.ai.planner import rewards as PR, mcts as PM
from oracle import rewards_ref as RW, mcts_ref as M
from oracle.philox import philox4x32_10


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 4096), st.integers(0, 2 ** 32 - 1))
def test_state_packing_round_trips_and_matches_oracle(A, seed):
    rng = np.random.default_rng(seed)
    applied = rng.random((3, A)) < rng.random()
    st_ = PR.pack_states(applied)
    NW, chunk, A_pad, nw = PR.layout(A)
    assert st_.shape == (3, nw) and (NW, chunk, A_pad, nw) == RW.layout(A)
    bits = RW.unpack_bits(st_)
    assert np.array_equal(bits[:, :A], applied) and bits[:, A:].all()          # padding bits are always "applied"
    assert np.array_equal(PR.empty_state(A), RW.empty_state(A))
    assert np.array_equal(PR.pack_states([list(np.nonzero(r)[0]) for r in applied], A), st_)


@settings(max_examples=40, deadline=None)
@given(st.integers(2, 300), st.integers(0, 2 ** 31 - 1))
def test_reward_is_additive_and_bounds_bracket(A, seed):
    rng = np.random.default_rng(seed)
    p = rng.random(A).astype(np.float32); size = (rng.random(A) * 5).astype(np.float32)
    cost = rng.choice([1.0, 10.0, 100.0], A).astype(np.float32)
    base = RW.empty_state(A)[None, :]
    s0 = float(RW.score(base, p, size, cost)[0])
    a = int(rng.integers(A))
    one = base.copy(); one[0, a >> 5] |= np.uint32(1) << np.uint32(a & 31)
    gain = float(RW.score(one, p, size, cost)[0]) - s0
    want = (2.0 * float(p[a]) - 1.0) * float(size[a]) - 0.1 * float(cost[a])
    assert abs(gain - want) <= 1e-4 * (1.0 + abs(s0))
    act = PR.Actions(p, size, cost)
    lo, inv = PR.reward_bounds(act)
    assert (lo, inv) == RW.reward_bounds(p, size, cost)
    for s in (s0, float(RW.score(one, p, size, cost)[0])):
        assert -1e-4 <= (s - float(lo)) * float(inv) <= 1 + 1e-4


@settings(max_examples=30, deadline=None)
@given(st.integers(0, 2 ** 32 - 1), st.integers(0, 2 ** 32 - 1), st.integers(0, 2 ** 32 - 1))
def test_philox_streams_differ_per_counter_and_key(c0, k0, k1):
    a = philox4x32_10(c0, 1, 2, 0, k0, k1); b = philox4x32_10((c0 + 1) & 0xFFFFFFFF, 1, 2, 0, k0, k1)
    c = philox4x32_10(c0, 1, 2, 0, k0 ^ 1, k1)
    assert tuple(map(int, a)) != tuple(map(int, b)) and tuple(map(int, a)) != tuple(map(int, c))
    assert tuple(map(int, a)) == tuple(map(int, philox4x32_10(c0, 1, 2, 0, k0, k1)))


@settings(max_examples=25, deadline=None)
@given(st.integers(2, 400), st.integers(0, 3000), st.integers(1, 9), st.integers(0, 2 ** 31 - 1))
def test_row_cuts_partition_rows_and_edges(N, E, parts, seed):
    rng = np.random.default_rng(seed)
    dst = np.sort(rng.integers(0, N, E)); rowptr = np.zeros(N + 1, np.int32); np.cumsum(np.bincount(dst, minlength=N), out=rowptr[1:])
    cuts = G.edge_balanced_row_cuts(rowptr, parts)
    assert cuts[0] == 0 and cuts[-1] == N and len(cuts) == parts + 1 and (np.diff(cuts) >= 0).all()
    assert int(np.diff(rowptr.astype(np.int64)[cuts]).sum()) == E


@settings(max_examples=15, deadline=None)
@given(st.integers(2, 40), st.integers(0, 2 ** 31 - 1))
def test_mcts_visit_accounting(A, seed):
    rng = np.random.default_rng(seed)
    p = rng.random(A).astype(np.float32); size = (0.1 + rng.random(A)).astype(np.float32); cost = np.ones(A, np.float32)
    R, D, T = 32, 6, 7
    r = M.search(p, size, cost, R=R, D=D, T=T, seed=int(seed % 1000))
    first = min(R, A)                                          # the stratified first pass covers min(R, A) children
    assert int(r["root_n"].sum()) == R + (T - 1) * R and int((r["root_n"] > 0).sum()) >= first
    assert 1 <= r["num_nodes"] <= T and r["best"] == M.best_child(r["root_n"], r["root_w"]) == PM.best_child(r["root_n"], r["root_w"])
    assert list(PM.ranked_children(r["root_n"], r["root_w"]))[0] == r["best"]
