"""Known-answer and property tests for the rewards / MCTS oracle (CPU)."""
import math

import numpy as np
import pytest

from oracle import mcts_ref as M
from oracle import rewards_ref as RW


def _actions(A, seed=2):
    rng = np.random.default_rng(seed)
    p = rng.beta(0.5, 0.5, A).astype(np.float32)
    size = rng.lognormal(np.log(2.0), 1.0, A).astype(np.float32)
    cost = rng.choice([1.0, 10.0, 100.0], A, p=[.9, .09, .01]).astype(np.float32)
    return p, size, cost


def test_reward_hand_computed():
    # threat-model.mdx:205-223 style candidates: revert (cost 1), kill (cost 10), restore (cost 100)
    p = np.array([0.95, 0.80, 1.0], np.float32)
    size = np.array([1.0, 0.0, 4.0], np.float32)
    cost = np.array([1.0, 10.0, 100.0], np.float32)
    st = np.stack([RW.empty_state(3)] * 3)
    st[1, 0] |= 0b001          # revert file
    st[2, 0] |= 0b101          # revert + restore
    s = RW.score(st, p, size, cost)
    # nothing undone: loss = .95*1 + .8*0 + 1*4 = 4.95
    assert s[0] == pytest.approx(-4.95, rel=1e-6)
    # revert file_1: loss = .05*1 + 4 = 4.05, downtime 1 -> -(4.05 + 0.1)
    assert s[1] == pytest.approx(-4.15, rel=1e-6)
    # + restore: loss = .05 + 0 = .05, downtime 101 -> -(0.05 + 10.1)
    assert s[2] == pytest.approx(-10.15, rel=1e-6)


@pytest.mark.parametrize("A", [1, 31, 100, 1024, 1500, 4096])
def test_reward_matches_float64(A):
    p, size, cost = _actions(A)
    rng = np.random.default_rng(A)
    applied = rng.random((16, A)) < 0.3
    NW, chunk, A_pad, nw = RW.layout(A)
    bits = np.ones((16, A_pad), bool); bits[:, :A] = applied
    st = RW.pack_bits(bits)
    got = RW.score(st, p, size, cost)
    p64, s64, c64 = p.astype(np.float64), size.astype(np.float64), cost.astype(np.float64)
    want = -(np.where(applied, (1 - p64) * s64, p64 * s64).sum(1) + 0.1 * (applied * c64).sum(1))
    np.testing.assert_allclose(got, want, rtol=2e-6)
    assert np.array_equal(RW.unpack_bits(st)[:, :A], applied)


def test_tree_sum_order():
    v = np.array([1e8, 1.0, -1e8, 1.0], np.float32)
    # adjacent pairs: (1e8+1) + (-1e8+1) = 1e8 + -1e8 = 0 in fp32
    assert M.tree_sum(v) == np.float32(np.float32(v[0] + v[1]) + np.float32(v[2] + v[3]))


def test_bounds_bracket_all_states():
    p, size, cost = _actions(10)
    lo, inv = RW.reward_bounds(p, size, cost)
    allst = RW.pack_bits(np.concatenate([
        ((np.arange(1024)[:, None] >> np.arange(10)) & 1).astype(bool), np.ones((1024, 1014), bool)], axis=1))
    s = RW.score(allst, p, size, cost)
    val = (s - lo) * inv
    assert val.min() >= -1e-5 and val.max() <= 1 + 1e-5
    assert val.max() > 0.999 and val.min() < 1e-3


def test_three_action_exhaustive_tree():
    """3 actions, R=4, D=3: every structural quantity can be checked by hand."""
    p = np.array([0.9, 0.1, 0.6], np.float32)
    size = np.array([2.0, 3.0, 1.0], np.float32)
    cost = np.array([1.0, 1.0, 1.0], np.float32)
    out = M.search(p, size, cost, R=4, D=3, T=1, seed=0, return_tree=True)
    # iteration 0: leaf = root, L0 = 3, stratified first moves: r=0,1,2,3 -> actions 0,1,2,0
    assert out["root_n"].tolist() == [2, 1, 1]
    assert out["tree"].visits[0] == 1 and out["num_nodes"] == 1
    # D=3 == A: every rollout ends with all actions applied -> identical terminal value
    full = RW.pack_bits(np.ones((1, 1024), bool))
    v = (RW.score(full, p, size, cost)[0] - out["lo"]) * out["inv_range"]
    assert out["root_w"][1] == np.float32(v) and out["root_w"][0] == np.float32(np.float32(v) + np.float32(v))
    out = M.search(p, size, cost, R=4, D=3, T=5, seed=0, return_tree=True)
    t = out["tree"]
    assert t.visits[0] == 5 and out["num_nodes"] == 5
    assert int(out["root_n"].sum()) == 4 + 4 * 4          # stratified + 4 descents x R
    # D == A: all terminal states are identical, so every Q is equal and UCT just balances visits
    q = out["root_w"] / out["root_n"]
    assert np.allclose(q, q[0], rtol=1e-6)
    assert out["root_n"].max() - out["root_n"].min() <= 4


def test_search_finds_best_gain_action_and_is_deterministic():
    p, size, cost = _actions(100)
    a = M.search(p, size, cost, R=256, D=20, T=40, seed=5)
    b = M.search(p, size, cost, R=256, D=20, T=40, seed=5)
    assert np.array_equal(a["root_n"], b["root_n"]) and np.array_equal(a["root_w"].view(np.uint32), b["root_w"].view(np.uint32))
    gain = (2 * p - 1) * size - 0.1 * cost
    assert (gain > gain[a["best"]]).sum() <= 2
    c = M.search(p, size, cost, R=256, D=20, T=40, seed=6)
    assert not np.array_equal(a["root_w"], c["root_w"])


def test_root_state_respected():
    p, size, cost = _actions(40)
    root = RW.empty_state(40); root[0] |= np.uint32(0b1011)
    out = M.search(p, size, cost, R=64, D=10, T=8, seed=1, root_state=root)
    assert out["root_n"][[0, 1, 3]].sum() == 0 and out["root_n"][2] > 0


def test_ln_table():
    t = M.ln_table(5, 4096)
    assert t[1] == np.float32(math.log(4096.0)) and t[5] == np.float32(math.log(5 * 4096.0))
