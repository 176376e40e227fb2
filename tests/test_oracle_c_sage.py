"""Second witness for the GraphSAGE-T oracle: oracle/c/sage_oracle.c (plain C + OpenMP) against oracle/sage_ref.py
(PyTorch index_select / index_add_ / addmm) and against an fp64 evaluation.

Two independent restatements of the frozen spec must agree to fp32 rounding: the aggregate
(same sequential per-row order in both) to ~1e-7, a whole layer to ~5e-6 of the output rms (max over all elements) (the 2F-term dot product
is associated differently by MKL and by the C loop), the 3-layer forward to ~1e-5 of the rms.  Both must sit equally
close to the fp64 evaluation -- neither is "the" truth."""
import numpy as np
import pytest
import torch

from nerrf_b200.graph import synthetic_graph, csr_from_edges
from oracle import c_sage, sage_ref as S

t = lambda a: torch.from_numpy(np.ascontiguousarray(a))


def _rel_rms(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.sqrt((b ** 2).mean()), 1e-30))


def test_hand_computed_toy():
    rowptr = np.array([0, 0, 2, 4, 5], np.int32)
    col = np.array([0, 2, 1, 3, 0], np.int32)
    ew = np.array([1.0, 3.0, 0.5, 0.5, 2.0], np.float32)
    x = np.array([[1.0, 0.0], [0.0, 1.0], [2.0, 2.0], [4.0, -4.0]], np.float32)
    m = c_sage.aggregate(x, rowptr, col, ew)
    want = np.array([[0, 0], [7 / 4, 6 / 4], [2.0, -1.5], [1.0, 0.0]], np.float32)
    assert np.allclose(m, want, atol=1e-7)
    W = np.array([[1.0, 0.0], [0.0, 1.0], [1.0, -1.0], [0.5, 2.0]], np.float32)
    b = np.array([0.1, -10.0], np.float32)
    h = c_sage.layer(x, rowptr, col, ew, W, b)
    assert np.allclose(h[1], [2.6, 0.0], atol=1e-6) and np.allclose(h[0], [1.1, 0.0], atol=1e-6)
    hl = c_sage.layer(x, rowptr, col, ew, W, b, relu=False)
    assert np.allclose(hl[1], [2.6, -7.75], atol=1e-6)


@pytest.mark.parametrize("F", [32, 64, 128])
@pytest.mark.parametrize("hub", ["src", "dst"])
def test_aggregate_and_layer_match_torch_oracle(F, hub):
    g = synthetic_graph(N=3000, E=40000, seed=11 + F, hub=hub, f_in=F)
    m_c = c_sage.aggregate(g.x, g.rowptr, g.col, g.ew)
    m_t = S.aggregate(t(g.x), t(g.rowptr), t(g.col), t(g.ew)).numpy()
    # same per-row sequential order, but one side may fuse the multiply-add: a few ulp
    assert _rel_rms(m_c, m_t) < 2e-6
    P = S.make_params(F, 128, 1, seed=5)
    W, b = P["layers"][0]
    h_c = c_sage.layer(g.x, g.rowptr, g.col, g.ew, W, b)
    h_t = S.layer(t(g.x), t(g.rowptr), t(g.col), t(g.ew), W, b).numpy()
    h_64 = S.layer(t(g.x), t(g.rowptr), t(g.col), t(g.ew), W, b, dtype=torch.float64).numpy()
    assert _rel_rms(h_c, h_t) < 1e-5
    # neither witness is closer to the exact result than fp32 rounding allows, and both are that close
    assert _rel_rms(h_c, h_64) < 1e-5 and _rel_rms(h_t, h_64) < 1e-5


def test_forward_matches_torch_oracle_and_fp64():
    g = synthetic_graph(N=20000, E=200000, seed=3)
    P = S.make_params(32, 128, 3, seed=1)
    h_c, s_c = c_sage.forward(P, g.x, g.rowptr, g.col, g.ew)
    h_t, s_t = S.forward(P, t(g.x), t(g.rowptr), t(g.col), t(g.ew))
    h_64, s_64 = S.forward(P, t(g.x), t(g.rowptr), t(g.col), t(g.ew), dtype=torch.float64)
    assert _rel_rms(h_c, h_t.numpy()) < 3e-5
    assert _rel_rms(h_c, h_64.numpy()) < 3e-5 and _rel_rms(h_t.numpy(), h_64.numpy()) < 3e-5
    assert np.abs(s_c - s_t.numpy()).max() < 2e-6 and np.abs(s_c - s_64.numpy()).max() < 2e-6
    # the ranking the pipeline consumes: top-64 anomalous nodes identical between the two witnesses
    assert np.array_equal(np.argsort(-s_c, kind="stable")[:64], np.argsort(-s_t.numpy(), kind="stable")[:64])


def test_row_range_int64_rowptr_and_empty_rows():
    rng = np.random.default_rng(0)
    N, E = 500, 3000
    dst = rng.integers(100, 300, E)                      # rows < 100 and >= 300 are isolated
    src = rng.integers(0, N, E)
    tt = rng.random(E).astype(np.float32) * 60
    conf = np.ones(E, np.float32)
    rowptr, col, ew = csr_from_edges(src, dst, tt, conf, N)
    x = rng.standard_normal((N, 32)).astype(np.float32)
    P = S.make_params(32, 128, 1, seed=2)
    W, b = P["layers"][0]
    full = c_sage.layer(x, rowptr, col, ew, W, b)
    part = c_sage.layer(x, rowptr.astype(np.int64), col, ew, W, b, row_begin=97, row_end=303)
    assert np.array_equal(full[97:303], part)            # row range + int64 rowptr: bit-identical
    m = c_sage.aggregate(x, rowptr, col, ew)
    assert not m[:100].any() and not m[300:].any()       # isolated node -> 0
    e = c_sage.layer(x, rowptr, col, ew, W, b, row_begin=10, row_end=10)
    assert e.shape == (0, 128)


def test_out_of_range_source_is_an_error():
    rowptr = np.array([0, 1], np.int32)
    with pytest.raises(ValueError):
        c_sage.aggregate(np.ones((1, 32), np.float32), rowptr, np.array([5], np.int32), np.ones(1, np.float32))
