"""Fleet-scale streamed traces (nerrf_b200/stream.py, BASELINE config 5): the columnar generator, the sliding window and
the candidate-only sequence builder.  Host code: runs on the CPU box."""
import numpy as np

from nerrf_b200 import graph as G, ingest, stream


def test_fleet_trace_is_the_m1_schema_replicated_with_renaming():
    cols, enc = stream.fleet_columns(60, 4, seed=5)
    assert len(enc) == 4 * 45 and all(p.endswith(".lockbit3") for p in enc)
    g = ingest.graph_from_columns(cols, observable=True)
    names = g.meta["names"]
    assert {names[i] for i in np.nonzero(g.meta["label"])[0]} == enc          # labels == the files a correct plan reverts
    kind = np.asarray(g.meta["node_kind"])
    assert int((kind == 1).sum()) == 60                                       # one process node per fleet member
    assert g.num_nodes > 60 * 95                                              # every process has its own 95 files
    # shared directories stay shared: /app/uploads is touched by every process
    up = names.index("/app/uploads")
    assert int(np.diff(g.rowptr)[up]) >= 60
    # same seed -> same bytes; per-event loader agrees with the columnar constructor on the fleet trace as well
    cols2, enc2 = stream.fleet_columns(60, 4, seed=5)
    assert enc2 == enc and np.array_equal(cols2.strings["path"][1], cols.strings["path"][1]) and np.array_equal(cols2.pid, cols.pid)
    small, _ = stream.fleet_columns(6, 2, seed=1)
    want = G.graph_from_events(ingest.events_from_columns(small), observable=True)
    got = ingest.graph_from_columns(small, observable=True)
    assert np.array_equal(got.rowptr, want.rowptr) and np.array_equal(got.col, want.col) and got.meta["names"] == want.meta["names"]
    assert np.array_equal(got.x.view(np.uint32), want.x.view(np.uint32))


def test_sliding_window_selects_by_time_and_keeps_columns_consistent():
    cols, _ = stream.fleet_columns(20, 3, seed=2)
    ts = cols.timestamp
    t0 = float(ts.min())
    w = stream.window(cols, t0 + 30.0, t0 + 90.0)
    keep = (ts > t0 + 30.0) & (ts <= t0 + 90.0)
    assert w.n == int(keep.sum()) and 0 < w.n < cols.n
    assert np.array_equal(w.pid, cols.pid[keep]) and np.array_equal(w.bytes, cols.bytes[keep])
    idx = np.nonzero(keep)[0]
    for k in (0, w.n // 2, w.n - 1):
        assert w.text("path", k) == cols.text("path", int(idx[k])) and w.text("syscall", k) == cols.text("syscall", int(idx[k]))
    empty = stream.window(cols, t0 - 10.0, t0 - 5.0)
    assert empty.n == 0 and empty.strings["path"][0].shape == (1,)
    # the window graph only knows what happened inside the window
    g = ingest.graph_from_columns(w, observable=True, window=60.0)
    assert g.num_nodes < ingest.graph_from_columns(cols, observable=True).num_nodes


def test_candidate_only_sequences_equal_the_full_builder():
    cols, _ = stream.fleet_columns(15, 3, seed=3)
    full = ingest.sequences_from_columns(cols, observable=True)
    pick = full[2][::5]
    sub = ingest.sequences_from_columns(cols, observable=True, only_nodes=pick)
    assert np.array_equal(sub[2], pick) and np.array_equal(sub[1], full[1][::5])
    assert np.array_equal(sub[0].view(np.uint32), full[0][::5].view(np.uint32))
    none = ingest.sequences_from_columns(cols, observable=True, only_nodes=np.zeros(0, np.int64))
    assert none[0].shape[0] == 0 and none[2].shape == (0,)


def test_name_hash_and_lazy_names():
    """stream.name_hash is a pure function of the name (the device kernel nerrf_trace_name_hash computes the same value --
    checked on the GPU in tests/test_gpu_stream.py); LazyNames decodes exactly the interning's naming events."""
    from nerrf_b200 import ingest
    assert stream.name_hash("/app/uploads/a.dat") == stream.name_hash(b"/app/uploads/a.dat")
    assert stream.name_hash("/app/uploads/a.dat") != stream.name_hash("/app/uploads/a.dat.lockbit3")
    assert stream.name_hash("pid:4242") != stream.name_hash("4242") and -(1 << 63) <= stream.name_hash("pid:4242") < (1 << 63)
    assert stream.name_hash("") != stream.name_hash("\0")                     # the length seeds the hash
    cols, _ = stream.fleet_columns(6, 2, seed=1)
    order = np.argsort(cols.timestamp, kind="stable")
    _, _, _, kind, name_event, name_which = ingest.intern_nodes(cols, order)
    want = ingest._node_names(cols, name_event, name_which)
    lazy = stream.LazyNames(cols, name_event, name_which)
    assert len(lazy) == len(want) and all(lazy[i] == want[i] for i in range(0, len(want), 7))
    assert len({stream.name_hash(n) for n in want}) == len(set(want))          # no collisions among the fleet's names
