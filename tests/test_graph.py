"""Host-side graph construction: synthetic generator, trace -> CSR, sharding cuts (CPU)."""
import numpy as np
import pytest

from nerrf_b200 import graph as G
from nerrf_b200 import trace_sim


def _check_csr(g):
    rp = g.rowptr.astype(np.int64)
    assert rp[0] == 0 and rp[-1] == g.num_edges and (np.diff(rp) >= 0).all()
    assert g.col.dtype == np.int32 and g.col.min() >= 0 and g.col.max() < g.num_nodes
    assert g.ew.dtype == np.float32 and (g.ew > 0).all() and (g.ew <= 1.0 + 1e-6).all()
    assert g.x.shape == (g.num_nodes, G.F_IN) and g.x.dtype == np.float32


def test_synthetic_generator_shape_and_skew():
    g = G.synthetic_graph(N=2000, E=20000, seed=20250115)
    _check_csr(g)
    # cubic skew: half of all edges come from the first N/8 sources
    assert 0.45 < (g.col < 250).mean() < 0.55
    h = G.synthetic_graph(N=2000, E=20000, seed=20250115)
    assert np.array_equal(g.col, h.col) and np.array_equal(g.ew, h.ew) and np.array_equal(g.x, h.x)
    d = G.synthetic_graph(N=2000, E=20000, hub="dst")
    assert np.diff(d.rowptr).max() > 5 * np.diff(g.rowptr).max()


def test_weights_are_time_decayed_and_rows_time_sorted():
    src = np.array([0, 1, 2, 0]); dst = np.array([3, 3, 3, 1])
    t = np.array([50.0, 10.0, 30.0, 60.0], np.float32); conf = np.ones(4, np.float32)
    rowptr, col, ew = G.csr_from_edges(src, dst, t, conf, 4)
    assert rowptr.tolist() == [0, 0, 1, 1, 4]
    assert col.tolist() == [0, 1, 2, 0]                      # row 3 sorted by time: t=10,30,50
    assert ew[0] == pytest.approx(1.0) and ew[1] == pytest.approx(np.exp(-50 / 30), rel=1e-6)
    assert ew[1] < ew[2] < ew[3]


def test_trace_to_graph_lockbit_sim():
    ev = trace_sim.lockbit_trace(n_files=12, seed=4)
    g = G.graph_from_events(ev)
    _check_csr(g)
    names = g.meta["names"]; y = g.meta["label"]
    enc = [n for n, lab in zip(names, y) if lab == 1]
    assert len(enc) == 12 and all(n.endswith(".lockbit3") for n in enc)
    assert sum(1 for k in g.meta["node_kind"] if k == 1) == 1            # one process node
    # every file node has the process as in-neighbour and vice versa
    p = list(g.meta["node_kind"]).index(1)
    rp = g.rowptr
    for v in range(g.num_nodes):
        if v != p:
            assert p in g.col[rp[v]:rp[v + 1]]
    # unordered (lossy, re-ordered) streams give the same graph
    rng = np.random.default_rng(0)
    ev2 = [ev[i] for i in rng.permutation(len(ev))]
    g2 = G.graph_from_events(ev2)
    assert np.array_equal(g.rowptr, g2.rowptr) and np.allclose(g.x, g2.x)


def test_replication_scales_nodes():
    ev = trace_sim.lockbit_trace(n_files=5, seed=1)
    g1 = G.graph_from_events(ev)
    g4 = G.graph_from_events(G.replicate_events(ev, 4))
    assert g4.num_nodes == 4 * g1.num_nodes and g4.num_edges == 4 * g1.num_edges


def test_edge_balanced_row_cuts():
    g = G.synthetic_graph(N=5000, E=60000, hub="dst")
    for parts in (1, 2, 3, 8):
        cuts = G.edge_balanced_row_cuts(g.rowptr, parts)
        assert cuts[0] == 0 and cuts[-1] == g.num_nodes and (np.diff(cuts) >= 0).all() and len(cuts) == parts + 1
        e = np.diff(g.rowptr.astype(np.int64)[cuts])
        assert e.sum() == g.num_edges
        if parts > 1:
            assert e.max() <= g.num_edges / parts + np.diff(g.rowptr).max()


def test_event_batch_adapter():
    class TS:                       # mimics google.protobuf.Timestamp
        def __init__(self, s): self.seconds, self.nanos = int(s), int((s % 1) * 1e9)
    class E:
        def __init__(self, ts, pid, syscall, path, nbytes, new_path=""):
            self.ts, self.pid, self.syscall, self.path, self.bytes, self.new_path = TS(ts), pid, syscall, path, nbytes, new_path
    class B: events = [E(1.5, 7, "openat", "/a/f.dat", 0), E(2.0, 7, "write", "/a/f.dat", 100),
                       E(2.5, 7, "rename", "/a/f.dat", 0, "/a/f.lockbit3")]
    ev = G.events_from_event_batch(B)
    g = G.graph_from_events(ev, merge_renames=False)
    _check_csr(g)
    assert g.num_nodes == 3        # pid, f.dat, f.lockbit3 (linked file<->file)
