"""Pin the numpy planner oracle against the independent C restatement (oracle/c/planner_oracle.c): bit-exact."""
import numpy as np
import pytest

from oracle import c_oracle, mcts_ref as M, rewards_ref as RW
from oracle.philox import philox4x32_10


def _actions(A, seed):
    rng = np.random.default_rng(seed)
    return (rng.beta(0.5, 0.5, A).astype(np.float32), rng.lognormal(np.log(2.0), 1.0, A).astype(np.float32),
            rng.choice([1.0, 10.0, 100.0], A, p=[.9, .09, .01]).astype(np.float32))


def test_philox_c_matches_numpy_and_random123_kat():
    assert c_oracle.philox(0, 0, 0, 0, 0, 0) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    for ctr in [(1, 2, 3, 4), (0xdeadbeef, 7, 0, 0xffffffff)]:
        assert c_oracle.philox(*ctr, 11, 12) == tuple(int(v) for v in philox4x32_10(*ctr, 11, 12))


@pytest.mark.parametrize("A", [1, 33, 100, 1024, 1500, 4096])
def test_score_bit_exact(A):
    p, size, cost = _actions(A, A)
    rng = np.random.default_rng(A + 1)
    bits = np.ones((19, RW.layout(A)[2]), bool); bits[:, :A] = rng.random((19, A)) < 0.4
    st = RW.pack_bits(bits)
    assert np.array_equal(c_oracle.score(st, p, size, cost).view(np.uint32), RW.score(st, p, size, cost).view(np.uint32))


@pytest.mark.parametrize("A,R,D,T,seed", [(3, 4, 3, 6, 0), (100, 256, 20, 30, 5), (40, 64, 10, 25, 1), (1024, 1024, 50, 4, 0),
                                          (1500, 256, 30, 4, 7), (10, 2048, 50, 10, 3), (64, 1, 5, 20, 2)])
def test_mcts_bit_exact_numpy_vs_c(A, R, D, T, seed):
    p, size, cost = _actions(A, seed + 10)
    a = M.search(p, size, cost, R=R, D=D, T=T, seed=seed)
    b = c_oracle.search(p, size, cost, R=R, D=D, T=T, seed=seed)
    assert np.array_equal(a["root_n"], b["root_n"]) and np.array_equal(a["root_w"].view(np.uint32), b["root_w"].view(np.uint32))
    assert a["best"] == b["best"] and a["num_nodes"] == b["num_nodes"]


def test_c_oracle_is_thread_count_independent_and_handles_root_state():
    p, size, cost = _actions(200, 3)
    root = RW.empty_state(200); root[2] |= np.uint32(0xFF00)
    a = c_oracle.search(p, size, cost, R=512, D=25, T=12, seed=4, root_state=root, threads=1)
    b = c_oracle.search(p, size, cost, R=512, D=25, T=12, seed=4, root_state=root, threads=4)
    w = M.search(p, size, cost, R=512, D=25, T=12, seed=4, root_state=root)
    for x in (a, b):
        assert np.array_equal(x["root_n"], w["root_n"]) and np.array_equal(x["root_w"].view(np.uint32), w["root_w"].view(np.uint32))
    assert a["root_n"][72:80].sum() == 0


# ---------------------------------------------------------------- planner spec v1: guards (non-separable reward)
def _guarded(A, n_kill, seed):
    """First n_kill actions are "kill process" candidates (size 0, cost 10); every other action is a file reversion
    guarded by one of them (or by none)."""
    rng = np.random.default_rng(seed)
    p, size, cost = _actions(A, seed)
    size[:n_kill] = 0.0; cost[:n_kill] = 10.0; cost[n_kill:] = 1.0
    guard = rng.integers(-1, n_kill, A).astype(np.int32); guard[:n_kill] = -1; guard[-1] = 0
    p[-1] = max(p[-1], 0.5); size[-1] = max(size[-1], 1.0); p[0] = max(p[0], 0.5)
    return p, size, cost, guard


def test_guard_term_hand_computed():
    # action 0 = kill (p .8, cost 10); action 1 = revert a 4 MB file, p = .9, guarded by 0
    p = np.array([0.8, 0.9], np.float32); size = np.array([0.0, 4.0], np.float32); cost = np.array([10.0, 1.0], np.float32)
    guard = np.array([-1, 0], np.int32)
    bits = np.ones((4, 1024), bool); bits[:, :2] = [[0, 0], [0, 1], [1, 0], [1, 1]]
    st = RW.pack_bits(bits)
    got = RW.score(st, p, size, cost, guard).astype(np.float64)
    u1, v1 = 0.9 * 4, 0.1 * 4
    want = [-(u1), -(v1 + 0.8 * u1 + 0.1 * 1), -(u1 + 0.1 * 10), -(v1 + 0.1 * 11)]
    assert np.allclose(got, want, rtol=1e-6)
    # the revert alone is worth LESS than doing nothing-but-kill-plus-revert: the value of an action depends on another
    assert got[3] > got[1] and got[3] > got[0] and got[2] < got[0]
    assert np.array_equal(RW.score(st, p, size, cost, None).view(np.uint32), RW.score(st, p, size, cost).view(np.uint32))


@pytest.mark.parametrize("A,n_kill", [(2, 1), (40, 3), (1024, 32), (1500, 7), (4096, 32)])
def test_guarded_score_bit_exact_numpy_vs_c(A, n_kill):
    p, size, cost, guard = _guarded(A, n_kill, A)
    rng = np.random.default_rng(A + 1)
    bits = np.ones((23, RW.layout(A)[2]), bool); bits[:, :A] = rng.random((23, A)) < 0.5
    st = RW.pack_bits(bits)
    a = RW.score(st, p, size, cost, guard); b = c_oracle.score(st, p, size, cost, guard)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert not np.array_equal(a, RW.score(st, p, size, cost))           # the term is really there


@pytest.mark.parametrize("A,n_kill,R,D,T,seed", [(12, 2, 64, 8, 30, 1), (100, 5, 256, 20, 20, 2), (1024, 32, 512, 30, 4, 3)])
def test_guarded_mcts_bit_exact_numpy_vs_c(A, n_kill, R, D, T, seed):
    p, size, cost, guard = _guarded(A, n_kill, seed)
    a = M.search(p, size, cost, R=R, D=D, T=T, seed=seed, guard=guard)
    b = c_oracle.search(p, size, cost, R=R, D=D, T=T, seed=seed, guard=guard)
    assert np.array_equal(a["root_n"], b["root_n"]) and np.array_equal(a["root_w"].view(np.uint32), b["root_w"].view(np.uint32))
    assert a["best"] == b["best"] and a["num_nodes"] == b["num_nodes"]


def test_threshold_rule_is_provably_wrong_under_spec_v1():
    """with the separable reward the optimum is a per-action threshold, so search was decorative.
    Under spec v1 it is not: 1 kill action + 8 guarded reversions.  Every action ALONE lowers the reward (the threshold
    rule therefore applies nothing); the exhaustive optimum applies all nine."""
    p = np.array([0.9] + [0.9] * 8, np.float32)
    size = np.array([0.0] + [1.0] * 8, np.float32)
    cost = np.array([10.0] + [1.0] * 8, np.float32)
    guard = np.array([-1] + [0] * 8, np.int32)
    A = 9
    bits = np.ones((1 << A, 1024), bool)
    bits[:, :A] = (np.arange(1 << A)[:, None] >> np.arange(A)) & 1
    sc = RW.score(RW.pack_bits(bits), p, size, cost, guard)
    empty = sc[0]
    singles = sc[[1 << a for a in range(A)]]
    assert (singles < empty).all()                                     # threshold rule: apply nothing
    best = int(np.argmax(sc))
    assert best == (1 << A) - 1 and sc[best] > empty + 3.0             # the real optimum: kill + revert everything
