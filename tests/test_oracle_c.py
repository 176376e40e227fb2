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
