"""GPU parity: device graph constructor (nerrf_graph_build_csr) vs the host constructor graph.csr_from_edges.
rowptr / col are index work: bit-exact.  ew = conf * exp(.): fp32, |err| <= 6 ulp of the host value (numpy's and
CUDA's expf are both faithfully rounded to within 1-2 ulp, not correctly rounded)."""
import numpy as np
import pytest
import torch

from nerrf_b200 import graph as G
from nerrf_b200._lib import NerrfError

pytestmark = pytest.mark.gpu


def _device_csr(src, dst, t, conf, N, **kw):
    c = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a.astype(dt))).cuda()
    rp, col, ew = G.build_csr_device(c(src, np.int32), c(dst, np.int32), c(t, np.float32), c(conf, np.float32), N, **kw)
    return rp.cpu().numpy(), col.cpu().numpy(), ew.cpu().numpy()


def _check(src, dst, t, conf, N, **kw):
    t = t.astype(np.float32); conf = conf.astype(np.float32)
    rp_w, col_w, ew_w = G.csr_from_edges(src.astype(np.int64), dst.astype(np.int64), t, conf, N, **kw)
    rp, col, ew = _device_csr(src, dst, t, conf, N, **kw)
    assert rp.dtype == rp_w.dtype
    assert np.array_equal(rp, rp_w)
    assert np.array_equal(col, col_w)
    assert np.all(np.abs(ew - ew_w) <= 6 * np.spacing(np.abs(ew_w)))
    return rp, col, ew


@pytest.mark.parametrize("N,E,seed", [(1, 5, 0), (7, 0, 1), (50, 400, 2), (1000, 20_000, 3), (100_000, 1_000_000, 4)])
def test_random_edge_lists(N, E, seed):
    rng = np.random.default_rng(seed)
    src = rng.integers(0, N, E); dst = rng.integers(0, N, E)
    t = rng.uniform(0, 60, E); conf = rng.uniform(0.5, 1.0, E)
    if E == 0:
        rp, col, ew = _device_csr(src, dst, t, conf, N)
        assert np.array_equal(rp, np.zeros(N + 1, np.int32)) and col.size == 0 and ew.size == 0
    else:
        _check(src, dst, t, conf, N)


def test_ties_keep_input_order_and_signed_zero_times():
    # many equal (dst, t) pairs: the stable order is the input order; -0.0 and +0.0 are the same time
    rng = np.random.default_rng(5)
    E, N = 5000, 40
    src = rng.integers(0, N, E); dst = rng.integers(0, 8, E)
    t = rng.integers(-2, 3, E).astype(np.float32) * np.float32(0.5)
    t[rng.random(E) < 0.1] = np.float32(-0.0)
    _check(src, dst, t, np.ones(E), N, t_ref=1.0)


def test_sparse_destinations_long_empty_runs():
    # edges land on a handful of far-apart rows: long empty-row gaps before, between and after
    N = 300_000
    dst = np.repeat(np.array([17, 18, 120_000, 299_000]), 300)
    rng = np.random.default_rng(6)
    src = rng.integers(0, N, dst.size); t = rng.uniform(0, 60, dst.size)
    rp, _, _ = _check(src, dst, t, np.ones(dst.size), N)
    assert rp[0] == 0 and rp[17] == 0 and rp[18] == 300 and rp[-1] == dst.size


def test_hub_destination_and_cfg2_shape_roundtrip_through_forward():
    # constructor output feeds GraphSAGE_T.forward directly; the result equals the forward on the host-built CSR
    from nerrf_b200.ai.models import GraphSAGE_T
    g_src = G.synthetic_graph(N=20_000, E=200_000, seed=7, hub="dst")
    rows = np.repeat(np.arange(g_src.num_nodes), np.diff(g_src.rowptr))
    rng = np.random.default_rng(8)
    shuffle = rng.permutation(rows.size)                      # undo the sort: a raw, unordered edge list
    src = g_src.col[shuffle]; dst = rows[shuffle]
    t = rng.uniform(0, 60, rows.size).astype(np.float32); conf = rng.uniform(0.5, 1, rows.size).astype(np.float32)
    rp_w, col_w, ew_w = G.csr_from_edges(src.astype(np.int64), dst.astype(np.int64), t, conf, g_src.num_nodes)
    c = lambda a: torch.from_numpy(a).cuda()
    rp, col, ew = G.build_csr_device(c(src.astype(np.int32)), c(dst.astype(np.int32)), c(t), c(conf), g_src.num_nodes)
    assert torch.equal(rp.cpu(), torch.from_numpy(rp_w)) and torch.equal(col.cpu(), torch.from_numpy(col_w))
    model = GraphSAGE_T(32, 128, 3).cuda()
    x = c(g_src.x)
    h_dev, s_dev = model(x, rp, col, ew)
    h_host, s_host = model(x, c(rp_w), c(col_w), c(ew_w))
    rms = float(h_host.pow(2).mean().sqrt())
    assert float((h_dev - h_host).abs().max()) <= 1e-5 * rms      # only the <=2-ulp weight difference separates them
    assert float((s_dev - s_host).abs().max()) <= 1e-5


def test_int64_rowptr_variant():
    rng = np.random.default_rng(9)
    N, E = 500, 9000
    src = rng.integers(0, N, E).astype(np.int32); dst = rng.integers(0, N, E).astype(np.int32)
    t = rng.uniform(0, 60, E).astype(np.float32); conf = np.ones(E, np.float32)
    c = lambda a: torch.from_numpy(a).cuda()
    rp64, col64, _ = G.build_csr_device(c(src), c(dst), c(t), c(conf), N, rowptr_dtype=torch.int64)
    rp32, col32, _ = G.build_csr_device(c(src), c(dst), c(t), c(conf), N)
    assert rp64.dtype == torch.int64 and torch.equal(rp64, rp32.long()) and torch.equal(col64, col32)


def test_out_of_range_vertex_is_an_error_not_a_crash():
    c = lambda a: torch.from_numpy(a).cuda()
    src = np.array([0, 1, 2], np.int32); t = np.zeros(3, np.float32); conf = np.ones(3, np.float32)
    for bad_dst in (np.array([0, 9, 1], np.int32), np.array([0, -1, 1], np.int32)):
        with pytest.raises(NerrfError, match="outside"):
            G.build_csr_device(c(src), c(bad_dst), c(t), c(conf), 3)
    with pytest.raises(NerrfError, match="outside"):
        G.build_csr_device(c(np.array([0, 7, 2], np.int32)), c(np.array([0, 1, 2], np.int32)), c(t), c(conf), 3)
    rp, col, _ = G.build_csr_device(c(src), c(np.array([2, 0, 2], np.int32)), c(t), c(conf), 3)    # still usable after
    assert rp.tolist() == [0, 1, 1, 3] and col.tolist() == [1, 0, 2]


@pytest.mark.parametrize("name", ["m0", "m1"])
def test_reference_trace_graphs_rebuild_from_shuffled_edges(name):
    # golden CSR of the reference's m0 / m1 traces (tests/golden/make_golden.py) -> shuffled edge list -> device
    # constructor gives the golden CSR back.  time = slot inside the golden CSR, so the (dst, t) order is the golden order.
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_m1_graph.npz"))
    rowptr, col, ew = z[name + "_rowptr"], z[name + "_col"], z[name + "_ew"]
    rows = np.repeat(np.arange(rowptr.size - 1), np.diff(rowptr))
    order = np.random.default_rng(10).permutation(col.size)
    slot = np.arange(col.size, dtype=np.float32)
    rp, c2, w2 = _device_csr(col[order], rows[order], slot[order], ew[order], rowptr.size - 1, t_ref=0.0, tau=1e30)
    assert np.array_equal(rp, rowptr) and np.array_equal(c2, col) and np.array_equal(w2, ew)   # exp(~0) == 1 exactly


def _features_close(dev, host):
    # x: counts / flags / ratios exact; log1p columns within 1 ulp of fp32 (CUDA's and the host's double log1p
    # can differ in the last bit of the double, which survives the fp32 rounding only at a tie)
    xd = dev.x.cpu().numpy()
    assert xd.shape == host.x.shape
    assert np.all(np.abs(xd - host.x) <= np.spacing(np.abs(host.x))), np.argwhere(xd != host.x)[:5]
    exact = [0, 1, 2, 14, 15, 16, 17, 18, 19] + list(range(20, 32))
    assert np.array_equal(xd[:, exact], host.x[:, exact])
    assert np.array_equal(dev.meta["label"], host.meta["label"])
    assert np.array_equal(dev.meta["size_mb"], host.meta["size_mb"])


@pytest.mark.parametrize("merge", [True, False])
def test_device_constructor_on_unordered_stream_with_rename_targets(merge):
    from nerrf_b200 import ingest
    rng = np.random.default_rng(12)
    ev = []
    for i in range(3000):
        kind = ["openat", "write", "rename", "file_encrypt_start", "file_encrypt_complete", "unlink"][int(rng.integers(6))]
        stem = "/app/uploads/f%d" % rng.integers(150)
        e = {"ts": (1_700_000_000 + int(rng.integers(0, 55)), int(rng.integers(0, 8)) * 125_000_000), "pid": int(rng.integers(1, 6)),
             "syscall": kind, "path": stem + [".dat", ".lockbit3", ""][int(rng.integers(3))], "bytes": int(rng.integers(0, 1 << 22))}
        if kind == "rename":
            e["new_path"] = "/app/uploads/f%d.lockbit3" % rng.integers(150)
        ev.append(e)
    cols = ingest.decode_event_batch(ingest.encode_event_batch(ev))
    host = ingest.graph_from_columns(cols, merge_renames=merge)
    dev = ingest.graph_from_columns(cols, merge_renames=merge, device="cuda")
    assert np.array_equal(dev.rowptr.cpu().numpy(), host.rowptr) and np.array_equal(dev.col.cpu().numpy(), host.col)
    assert np.all(np.abs(dev.ew.cpu().numpy() - host.ew) <= 6 * np.spacing(np.abs(host.ew)))
    _features_close(dev, host)
    assert dev.meta["names"] == host.meta["names"]


def test_node_features_reject_bad_columns():
    i32 = lambda *v: torch.tensor(v, dtype=torch.int32).cuda()
    args = dict(node_p=i32(0, 0), node_f=i32(1, 5), node_g=None, t=torch.tensor([0.0, 1.0], dtype=torch.float64).cuda(),
                event_slot=torch.tensor([4, 5], dtype=torch.uint8).cuda(), nbytes=torch.tensor([1, 2]).cuda(),
                path_flags=torch.zeros(2, dtype=torch.uint8).cuda(), node_kind=torch.tensor([1, 0], dtype=torch.int8).cuda(), window=60.0)
    with pytest.raises(NerrfError, match="outside"):
        G.node_features_device(**args)
    args["node_f"] = i32(1, 1)
    x, label, size_mb = G.node_features_device(**args)
    assert x.shape == (2, 32) and label.tolist() == [0, 0] and float(x[1, 5 + 4]) == pytest.approx(np.log1p(1.0))
    assert float(size_mb[1]) == pytest.approx(3.0 / 1e6) and float(x[0, 3]) == pytest.approx(np.log1p(2.0))


def test_wire_bytes_to_device_graph_to_plan():
    # EventBatch wire bytes -> columns -> graph with the CSR stage on the GPU -> the same graph as the host path,
    # and the pipeline runs on the device-resident graph without another upload
    from nerrf_b200 import ingest, pipeline, trace_sim
    from nerrf_b200.ai.models import GraphSAGE_T, lstm
    ev = trace_sim.lockbit_trace(n_files=20, seed=11, benign_files=15)
    cols = ingest.decode_event_batch(ingest.encode_event_batch(ev))
    host = ingest.graph_from_columns(cols)
    dev = ingest.graph_from_columns(cols, device="cuda")
    assert dev.rowptr.is_cuda and dev.x.is_cuda and dev.num_nodes == host.num_nodes
    assert np.array_equal(dev.rowptr.cpu().numpy(), host.rowptr) and np.array_equal(dev.col.cpu().numpy(), host.col)
    assert np.all(np.abs(dev.ew.cpu().numpy() - host.ew) <= 6 * np.spacing(np.abs(host.ew)))
    _features_close(dev, host)
    seq, lengths, nodes = pipeline.file_sequences(ingest.events_from_columns(cols), host)
    label = host.meta["label"].astype(bool)
    model = GraphSAGE_T(32, 128, 2).cuda(); seq_model = lstm.LSTMScorer().cuda()
    res = pipeline.run(dev, seq, lengths, nodes, model, seq_model, top_a=4096, confidence=np.where(label, 0.95, 0.05),
                       n_rollouts=512, depth=30, iterations=6)
    assert sorted(res.plan_nodes) == sorted(np.nonzero(label)[0].tolist())
