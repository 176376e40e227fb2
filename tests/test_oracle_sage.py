"""Known-answer tests for the GraphSAGE-T oracle (CPU)."""
import torch

from oracle import sage_ref as S


def _toy():
    # 4 nodes, 5 edges (src -> dst): 0->1, 2->1, 1->2, 3->2, 0->3 ; node 0 has no in-edges
    rowptr = torch.tensor([0, 0, 2, 4, 5], dtype=torch.int32)
    col = torch.tensor([0, 2, 1, 3, 0], dtype=torch.int32)
    ew = torch.tensor([1.0, 3.0, 0.5, 0.5, 2.0])
    x = torch.tensor([[1.0, 0.0], [0.0, 1.0], [2.0, 2.0], [4.0, -4.0]])
    return x, rowptr, col, ew


def test_aggregate_hand_computed():
    x, rowptr, col, ew = _toy()
    m = S.aggregate(x, rowptr, col, ew)
    want = torch.tensor([[0.0, 0.0],                                 # isolated -> 0
                         [(1 * 1 + 3 * 2) / 4, (0 + 3 * 2) / 4],     # (1*x0 + 3*x2)/4
                         [(0 + 0.5 * 4) / 1.0, (0.5 * 1 - 0.5 * 4) / 1.0],
                         [1.0, 0.0]])
    assert torch.allclose(m, want, atol=1e-7)


def test_layer_hand_computed():
    x, rowptr, col, ew = _toy()
    W = torch.tensor([[1.0, 0.0], [0.0, 1.0], [1.0, -1.0], [0.5, 2.0]])   # [2F=4, H=2]
    b = torch.tensor([0.1, -10.0])
    h = S.layer(x, rowptr, col, ew, W, b)
    # node 1: [0,1,1.75,1.5] @ W + b = [0+0+1.75+0.75+.1, 1-1.75+3-10] = [2.6, -7.75] -> relu
    assert torch.allclose(h[1], torch.tensor([2.6, 0.0]), atol=1e-6)
    # node 0 (isolated): [1,0,0,0] @ W + b = [1.1, -10]
    assert torch.allclose(h[0], torch.tensor([1.1, 0.0]), atol=1e-6)


def test_forward_fp32_vs_fp64_and_heads():
    from nerrf_b200.graph import synthetic_graph
    g = synthetic_graph(N=500, E=4000, seed=7)
    P = S.make_params(32, 128, 3, seed=1)
    t = lambda a: torch.from_numpy(a)
    h, sc, el = S.forward(P, t(g.x), t(g.rowptr), t(g.col), t(g.ew), edge_logits=True)
    h64, sc64, el64 = S.forward(P, t(g.x), t(g.rowptr), t(g.col), t(g.ew), edge_logits=True, dtype=torch.float64)
    assert h.shape == (500, 128) and sc.shape == (500,) and el.shape == (4000, 2)
    assert torch.allclose(h.double(), h64, rtol=1e-4, atol=1e-5)
    assert torch.allclose(sc.double(), sc64, atol=1e-5) and (sc > 0).all() and (sc < 1).all()
    assert torch.allclose(el.double(), el64, rtol=1e-4, atol=1e-4)


def test_row_range_matches_full():
    from nerrf_b200.graph import synthetic_graph
    g = synthetic_graph(N=300, E=2000, seed=3)
    t = lambda a: torch.from_numpy(a)
    P = S.make_params(32, 128, 1, seed=2)
    W, b = P["layers"][0]
    full = S.layer(t(g.x), t(g.rowptr), t(g.col), t(g.ew), W, b)
    part = S.layer(t(g.x), t(g.rowptr), t(g.col), t(g.ew), W, b, row_begin=100, row_end=250)
    assert torch.equal(full[100:250], part)
