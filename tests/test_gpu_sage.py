"""GraphSAGE-T CUDA path vs the oracle, through the C-ABI (needs a B200)."""
import numpy as np
import pytest
import torch

from nerrf_b200 import graph as G
from nerrf_b200.ai.models import GraphSAGE_T
from nerrf_b200.ai.models.graphsage_t import HostSession
from nerrf_b200 import _lib as L
from oracle import sage_ref as S
from gpu_util import dev_graph, cpu_graph, assert_close_fp32

pytestmark = pytest.mark.gpu
ALGOS = ["ffma", "umma", "umma2"]
# "umma2" (2-term bf16 split) is ~1e-5 relative to the tensor scale: within the north star's 1e-4 bound
# but not within the strict absolute floor used for the fp32-equivalent paths.
ATOL_RMS = {"ffma": 1e-5, "umma": 1e-5, "umma2": 1e-4, "auto": 1e-5}


@pytest.mark.parametrize("F", [32, 64, 128])
def test_aggregate_parity(F):
    g = G.synthetic_graph(N=3000, E=40000, seed=11, f_in=F)
    x, rp, col, ew = dev_graph(g)
    m = torch.empty(3000, F, device="cuda")
    L.check(L.lib().nerrf_sage_aggregate(L.ptr(x), L.ptr(rp), 0, L.ptr(col), L.ptr(ew), L.ptr(m), 3000, 0, 3000, F,
                                         L.current_stream_ptr()))
    want = S.aggregate(*cpu_graph(g))
    assert_close_fp32(m, want, what=f"aggregate F={F}")
    # isolated rows are exactly zero
    deg = np.diff(g.rowptr)
    assert (m.cpu()[torch.from_numpy(deg == 0)] == 0).all()


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("F", [32, 64, 128])
@pytest.mark.parametrize("hub", ["src", "dst"])
def test_layer_parity(algo, F, hub):
    g = G.synthetic_graph(N=5000, E=60000, seed=5, hub=hub, f_in=F)
    model = GraphSAGE_T(F, 128, 1, algo=algo).cuda()
    x, rp, col, ew = dev_graph(g)
    out = model.layer_forward(0, x, rp, col, ew)
    W, b = model.oracle_params()["layers"][0]
    want = S.layer(*cpu_graph(g), W, b)
    assert_close_fp32(out, want, atol_rms=ATOL_RMS[algo], what=f"layer {algo} F={F} hub={hub}")
    # int64 rowptr gives bit-identical output
    out64 = model.layer_forward(0, x, rp.long(), col, ew)
    assert torch.equal(out, out64)


@pytest.mark.parametrize("algo", ALGOS)
def test_layer_edge_cases(algo):
    model = GraphSAGE_T(32, 128, 1, algo=algo).cuda()
    W, b = model.oracle_params()["layers"][0]
    # (a) no edges at all, N not a multiple of the tile; (b) one node; (c) a single hub row with every edge
    for N, E, hubrow in [(77, 0, None), (1, 0, None), (200, 5000, 3)]:
        rng = np.random.default_rng(N)
        src = rng.integers(0, N, E); dst = np.full(E, hubrow if hubrow is not None else 0)
        t = rng.random(E).astype(np.float32) * 60; conf = np.ones(E, np.float32)
        rowptr, col, ew = G.csr_from_edges(src, dst, t, conf, N)
        g = G.TemporalGraph(rowptr, col, ew, rng.standard_normal((N, 32)).astype(np.float32), {})
        out = model.layer_forward(0, *dev_graph(g))
        assert_close_fp32(out, S.layer(*cpu_graph(g), W, b), atol_rms=ATOL_RMS[algo], what=f"edge case N={N} E={E}")
    # (d) row range: only [row_begin,row_end) is written
    g = G.synthetic_graph(N=1000, E=9000, seed=2)
    x, rp, col, ew = dev_graph(g)
    out = torch.full((1000, 128), -7.0, device="cuda")
    model.layer_forward(0, x, rp, col, ew, out=out, row_begin=130, row_end=901)
    full = model.layer_forward(0, x, rp, col, ew)
    assert torch.equal(out[130:901], full[130:901]) and (out[:130] == -7).all() and (out[901:] == -7).all()


@pytest.mark.parametrize("algo", ALGOS)
def test_forward_parity_and_indices(algo):
    """3-layer forward + heads; anomalous-node indices (top-k by score) must be bit-exact."""
    g = G.synthetic_graph(N=20000, E=200000, seed=20250115)
    model = GraphSAGE_T(32, 128, 3, algo=algo).cuda()
    h, sc, el = model(*dev_graph(g), return_edge_logits=True)
    hw, scw, elw = S.forward(model.oracle_params(), *cpu_graph(g), edge_logits=True)
    assert_close_fp32(h, hw, atol_rms=ATOL_RMS[algo], what="h")
    assert_close_fp32(sc, scw, atol_rms=ATOL_RMS[algo], what="node_score")
    assert_close_fp32(el, elw, rtol=1e-4, atol_rms=1e-4, what="edge_logit")
    k = 64
    top_gpu = torch.topk(sc.cpu(), k).indices
    top_ref = torch.topk(scw, k).indices
    margin = float((torch.sort(scw, descending=True).values[:k + 1].diff().abs()).min())
    assert torch.equal(top_gpu, top_ref), f"top-{k} anomalous node indices differ (min score margin {margin:.2e})"
    assert torch.equal(el.argmax(1).cpu(), elw.argmax(1)) or (elw[:, 0] - elw[:, 1]).abs().min() < 1e-5


def test_forward_is_deterministic():
    g = G.synthetic_graph(N=8000, E=90000, seed=1)
    model = GraphSAGE_T(32, 128, 3).cuda()
    a = model(*dev_graph(g)); b = model(*dev_graph(g))
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_trace_graph_parity():
    """cfg 1: toy LockBit trace (simulator schema) replicated x10 -> ~1k nodes, 2 layers."""
    from nerrf_b200 import trace_sim
    ev = G.replicate_events(trace_sim.lockbit_trace(n_files=45, seed=0, benign_files=50), 10)
    g = G.graph_from_events(ev)
    assert 900 <= g.num_nodes <= 1100
    model = GraphSAGE_T(32, 128, 2).cuda()
    h, sc = model(*dev_graph(g))
    hw, scw = S.forward(model.oracle_params(), *cpu_graph(g))
    assert_close_fp32(h, hw, what="toy h"); assert_close_fp32(sc, scw, what="toy score")


def test_host_session_matches_device_path():
    g = G.synthetic_graph(N=30000, E=250000, seed=9)
    model = GraphSAGE_T(32, 128, 3).cuda()
    h, sc = model(*dev_graph(g))
    sess = HostSession(model, 40000, 300000)
    pin = lambda a: torch.from_numpy(a).pin_memory()
    score = torch.empty(g.num_nodes).pin_memory(); hout = torch.empty(g.num_nodes, 128).pin_memory()
    sess.forward(pin(g.x), pin(g.rowptr), pin(g.col), pin(g.ew), score, hout)
    assert torch.equal(score, sc.cpu()) and torch.equal(hout, h.cpu())
    with pytest.raises(L.NerrfError):
        big = G.synthetic_graph(N=50000, E=1000, seed=1)
        sess.forward(pin(big.x), pin(big.rowptr), pin(big.col), pin(big.ew), torch.empty(50000).pin_memory())
    sess.close()


def test_host_session_pipelined_submit_wait():
    """nerrf_sage_session_submit_host / _wait: a stream of DIFFERENT graphs, two in flight; every ticket's outputs equal
    the device path's for its own graph (buffer sets never mix), in any wait order."""
    model = GraphSAGE_T(32, 128, 3).cuda()
    pin = lambda a: torch.from_numpy(a).pin_memory()
    graphs = [G.synthetic_graph(N=20000 + 3000 * i, E=1100000 + 50000 * i if i % 2 else 90000, seed=30 + i) for i in range(5)]
    want = [tuple(t.cpu() for t in model(*dev_graph(g))) for g in graphs]
    sess = HostSession(model, 40000, 1400000)
    host = [(pin(g.x), pin(g.rowptr), pin(g.col), pin(g.ew), torch.empty(g.num_nodes).pin_memory(),
             torch.empty(g.num_nodes, 128).pin_memory() if i % 2 == 0 else None) for i, g in enumerate(graphs)]
    tickets = []
    for i, hb in enumerate(host):                      # depth-2 pipeline: submit i, then wait i-1
        tickets.append(sess.submit(*hb))
        if i >= 1:
            sess.wait(tickets[i - 1])
    sess.wait(tickets[-1])
    sess.wait(tickets[0])                                # waiting an old ticket again is a no-op
    for i, hb in enumerate(host):
        assert torch.equal(hb[4], want[i][1]), f"scores of ticket {i}"
        if hb[5] is not None:
            assert torch.equal(hb[5], want[i][0]), f"embeddings of ticket {i}"
    # three submits without a wait: the third blocks on the oldest internally, results still right
    ts = [sess.submit(*host[i]) for i in (2, 3, 4)]
    for t in ts:
        sess.wait(t)
    assert torch.equal(host[4][4], want[4][1]) and torch.equal(host[2][4], want[2][1])
    with pytest.raises(L.NerrfError):
        sess.wait(10 ** 6)
    sess.close()


def _ranking_matches(sc_gpu, sc_ref, k, err):
    """Top-k anomalous-node indices.  Exact equality is required for the prefix of the oracle's ranking whose adjacent
    score margins exceed 20x the measured max score error (there a flip would be a real bug); past the first
    near-tie the two rankings may differ only by swaps of nodes whose ORACLE scores are within 4*err."""
    order_ref = np.argsort(-sc_ref, kind="stable")[:k + 1]
    order_gpu = np.argsort(-sc_gpu, kind="stable")[:k]
    margins = np.abs(np.diff(sc_ref[order_ref].astype(np.float64)))
    near = np.nonzero(margins < 20 * max(err, 1e-9))[0]
    safe = int(near[0]) if near.size else k
    assert np.array_equal(order_gpu[:safe], order_ref[:safe]), "ranking differs where the oracle's margin is large"
    assert np.abs(sc_ref[order_gpu].astype(np.float64) - sc_ref[order_ref[:k]].astype(np.float64)).max() <= 4 * err + 1e-12
    return safe


@pytest.mark.parametrize("family", ["hub_src", "hub_dst", "uniform"])
def test_full_size_forward_vs_oracle(family):
    """BASELINE cfg 2 at FULL size (1M nodes / 10M edges, 3 layers): the whole forward -- every element of h, every
    node score, the anomalous-node ranking -- against the oracle.  The oracle here is the C/OpenMP restatement
    (oracle/c/sage_oracle.c, itself pinned against oracle/sage_ref.py and fp64 in tests/test_oracle_c_sage.py), because
    it finishes the full graph in about a second on the box's host cores.  Families: the headline generator (hub
    sources), hub destinations (long rows: the pre-aggregation path), and uniform sources (no L2 help)."""
    from oracle import c_sage
    if family == "uniform":
        rng = np.random.Generator(np.random.PCG64(77))
        N, E = 1_000_000, 10_000_000
        rowptr, col, ew = G.csr_from_edges(rng.integers(0, N, E), rng.integers(0, N, E), (rng.random(E) * 60).astype(np.float32),
                                           (0.5 + 0.5 * rng.random(E)).astype(np.float32), N)
        g = G.TemporalGraph(rowptr, col, ew, np.random.Generator(np.random.PCG64(0)).standard_normal((N, 32), dtype=np.float32), {})
    else:
        g = G.synthetic_graph(hub="src" if family == "hub_src" else "dst")
    model = GraphSAGE_T(32, 128, 3).cuda()
    h, sc = model(*dev_graph(g))
    hw, scw = c_sage.forward(model.oracle_params(), g.x, g.rowptr, g.col, g.ew)
    assert_close_fp32(h, torch.from_numpy(hw), what=f"full-size h ({family})")
    err = float(np.abs(sc.cpu().numpy() - scw).max())
    assert err < 1e-5, f"node scores differ by {err}"
    safe = _ranking_matches(sc.cpu().numpy(), scw, 64, err)
    assert safe >= 4, f"only the first {safe} ranks have a margin above the tolerance -- pick another seed"
    # anomalous-node SET at the operating threshold is identical away from the threshold itself
    thr = float(np.sort(scw)[-1000])
    decided = np.abs(scw - thr) > 4 * err
    assert np.array_equal((sc.cpu().numpy() > thr)[decided], (scw > thr)[decided])


def test_full_size_properties():
    """BASELINE cfg 2 size (1M nodes / 10M edges): size-independent properties instead of the oracle.
    (1) constant features: the weighted mean of a constant is that constant for every non-isolated
        row; (2) linearity of the aggregate; (3) spot rows against an fp64 numpy evaluation."""
    g = G.synthetic_graph()              # 1M / 10M
    N = g.num_nodes
    x, rp, col, ew = dev_graph(g)
    deg = torch.from_numpy(np.diff(g.rowptr)).cuda()
    for F in (32, 128):
        ones = torch.ones(N, F, device="cuda")
        m = torch.empty(N, F, device="cuda")
        agg = lambda inp, outp: L.check(L.lib().nerrf_sage_aggregate(L.ptr(inp), L.ptr(rp), 0, L.ptr(col), L.ptr(ew), L.ptr(outp),
                                                                     N, 0, N, F, L.current_stream_ptr()))
        agg(ones, m)
        nz = deg > 0
        assert (m[nz] - 1).abs().max() < 1e-5 and (m[~nz] == 0).all()
        a = torch.randn(N, F, device="cuda"); b = torch.randn(N, F, device="cuda")
        ma = torch.empty_like(m); mb = torch.empty_like(m); mab = torch.empty_like(m)
        agg(a, ma); agg(b, mb); agg(2 * a - 3 * b, mab)
        assert (mab - (2 * ma - 3 * mb)).abs().max() < 1e-4
        rows = np.random.default_rng(0).integers(0, N, 64)
        a_cpu = a.cpu().numpy().astype(np.float64)
        for r in rows:
            e0, e1 = g.rowptr[r], g.rowptr[r + 1]
            if e1 > e0:
                w = g.ew[e0:e1].astype(np.float64)
                want = (a_cpu[g.col[e0:e1]] * w[:, None]).sum(0) / w.sum()
                np.testing.assert_allclose(ma[r].cpu().numpy(), want, rtol=1e-4, atol=1e-5)
    # full 3-layer forward: finite, ReLU-nonnegative, scores in (0,1), deterministic
    model = GraphSAGE_T(32, 128, 3).cuda()
    h, sc = model(x, rp, col, ew)
    h2, sc2 = model(x, rp, col, ew)
    assert torch.isfinite(h).all() and (h >= 0).all() and (sc > 0).all() and (sc < 1).all()
    assert torch.equal(h, h2) and torch.equal(sc, sc2)
    # spot-check 32 rows of the last layer against the oracle restricted to those rows' inputs:
    # recompute layer 3 for a row range with the oracle from the GPU's layer-2 output
    h1 = model.layer_forward(0, x, rp, col, ew); h2_ = model.layer_forward(1, h1, rp, col, ew)
    W, b = model.oracle_params()["layers"][2]
    want = S.layer(h2_.cpu(), torch.from_numpy(g.rowptr), torch.from_numpy(g.col), torch.from_numpy(g.ew), W, b,
                   row_begin=500000, row_end=500256)
    assert_close_fp32(h[500000:500256], want, what="full-size layer-3 rows")


@pytest.mark.parametrize("algo", ALGOS)
def test_fused_head_matches_separate_head(algo):
    g = G.synthetic_graph(N=7001, E=80000, seed=4, f_in=128)
    model = GraphSAGE_T(128, 128, 1, algo=algo).cuda()
    x, rp, col, ew = dev_graph(g)
    score = torch.full((7001,), -1.0, device="cuda")
    h = model.layer_forward(0, x, rp, col, ew, score_out=score)
    sep, _ = model.heads(h, rp, col)
    assert (score - sep).abs().max() < 1e-6
    # row range: scores outside [row_begin,row_end) untouched
    score2 = torch.full((7001,), -1.0, device="cuda")
    model.layer_forward(0, x, rp, col, ew, row_begin=100, row_end=6000, score_out=score2)
    assert torch.equal(score2[100:6000], score[100:6000]) and (score2[:100] == -1).all() and (score2[6000:] == -1).all()


@pytest.mark.parametrize("F", [32, 128])
def test_hub_rows_preaggregated_path_matches_inline_and_oracle(F):
    """Rows with > 128 in-edges take the chunk pre-aggregation path; same answer as the inline path and the oracle."""
    rng = np.random.default_rng(F)
    N, E = 4000, 120000
    dst = np.minimum((N * rng.random(E) ** 6).astype(np.int64), N - 1)          # a few rows with 10^3..10^4 in-edges
    src = rng.integers(0, N, E)
    t = (rng.random(E) * 60).astype(np.float32); conf = (0.5 + 0.5 * rng.random(E)).astype(np.float32)
    rowptr, col, ew = G.csr_from_edges(src, dst, t, conf, N)
    assert np.diff(rowptr).max() > 5000
    g = G.TemporalGraph(rowptr, col, ew, rng.standard_normal((N, F)).astype(np.float32), {})
    model = GraphSAGE_T(F, 128, 1, algo="umma").cuda()
    x, rp, c, w = dev_graph(g)
    out = model.layer_forward(0, x, rp, c, w)                                    # with the hub-row scratch
    W, b = model.oracle_params()["layers"][0]
    assert_close_fp32(out, S.layer(*cpu_graph(g), W, b), what="hub rows vs oracle")
    # inline path (no scratch) through the plain C-ABI entry
    out2 = torch.empty_like(out)
    L.check(L.lib().nerrf_sage_layer_fwd(L.ptr(x), L.ptr(rp), 0, L.ptr(c), L.ptr(w), L.ptr(model.weights[0]),
                                         L.ptr(model.biases[0]), L.ptr(out2), N, 0, N, F, 128, 1, 2, L.current_stream_ptr()))
    assert_close_fp32(out2, out, what="inline vs pre-aggregated")
    assert torch.equal(out, model.layer_forward(0, x, rp, c, w))                 # deterministic
