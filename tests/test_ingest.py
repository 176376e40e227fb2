"""EventBatch ingest (csrc/ingest.cu) pinned against the protobuf runtime, and the columnar graph constructor
pinned against the per-event host loader graph.graph_from_events.  All host code: runs on the CPU box."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from nerrf_b200 import graph as G, ingest, trace_sim
from nerrf_b200._lib import NerrfError
from proto_util import classes

pytestmark = pytest.mark.usefixtures("lib_built")


def _random_batch(rng, n, rich=False):
    Event, Batch = classes()
    b = Batch()
    syscalls = ["openat", "write", "rename", "unlink", "", "file_encrypt_start", "file_encrypt_complete", "file_created"]
    for _ in range(n):
        e = b.events.add()
        if rng.random() < 0.9:
            e.ts.seconds = int(rng.integers(-5, 2_000_000_000)); e.ts.nanos = int(rng.integers(0, 1_000_000_000))
        e.pid = int(rng.integers(0, 1 << 32)) if rng.random() < 0.5 else int(rng.integers(0, 6))
        e.tid = int(rng.integers(0, 1 << 32)) if rng.random() < 0.8 else 0
        e.comm = ["python3", "", "bash", "lockbit-sim"][int(rng.integers(4))]
        e.syscall = syscalls[int(rng.integers(len(syscalls)))]
        stem = ["/app/uploads/f%d" % rng.integers(20), "/tmp/x%d" % rng.integers(3), "/proc/%d/maps" % rng.integers(9),
                "/data/Readme_%d" % rng.integers(3), "/home/ü/ñ%d" % rng.integers(4), "", "/noext/dir.d/file"][int(rng.integers(7))]
        e.path = stem + ["", ".dat", ".lockbit3", ".txt", ".lockbit"][int(rng.integers(5))] if stem else ""
        if rng.random() < 0.3:
            e.new_path = "/app/uploads/f%d.lockbit3" % rng.integers(20)
        e.flags = int(rng.integers(0, 3))
        e.ret_val = int(rng.integers(-(1 << 62), 1 << 62)) if rng.random() < 0.5 else int(rng.integers(-40, 40))
        e.bytes = int(rng.integers(0, 1 << 63)) * 2 + 1 if rng.random() < 0.2 else int(rng.integers(0, 1 << 22))
        if rich:                                            # fields the graph does not use: must be skipped cleanly
            e.inode = str(rng.integers(1, 1 << 40)); e.mode = 0o644; e.uid = int(rng.integers(0, 1 << 40)); e.gid = 5
            e.dependencies.extend(["a", "/b/c"][: int(rng.integers(0, 3))])
    return b


def _assert_columns_equal(cols, batch):
    assert cols.n == len(batch.events)
    for i, e in enumerate(batch.events):
        assert (cols.ts_sec[i], cols.ts_nanos[i]) == (e.ts.seconds, e.ts.nanos)
        assert (cols.pid[i], cols.tid[i], cols.flags[i], cols.ret_val[i], cols.bytes[i]) == (e.pid, e.tid, e.flags, e.ret_val, e.bytes)
        assert cols.text("comm", i) == e.comm and cols.text("syscall", i) == e.syscall
        assert cols.text("path", i) == e.path and cols.text("new_path", i) == e.new_path
        assert cols.event_slot[i] == G._EVENT_SLOT.get(e.syscall, 7)
        want = (1 if ".lockbit" in e.path else 0) | (2 if ("README" in e.path.upper() or "RANSOM" in e.path.upper()) else 0) \
            | (4 if e.path.startswith(("/tmp", "/proc")) else 0) | (8 if e.path.endswith(".lockbit3") else 0)
        assert cols.path_flags[i] == want


@pytest.mark.parametrize("n,rich,seed", [(0, False, 0), (1, False, 1), (200, False, 2), (300, True, 3)])
def test_decoder_matches_protobuf_runtime(n, rich, seed):
    b = _random_batch(np.random.default_rng(seed), n, rich)
    _assert_columns_equal(ingest.decode_event_batch(b.SerializeToString()), b)


def test_concatenated_batches_are_one_batch():
    # the tracker sends one EventBatch per event (main.go:254); a buffered stream is their concatenation
    rng = np.random.default_rng(4)
    parts = [_random_batch(rng, k) for k in (3, 0, 5, 1)]
    _, Batch = classes()
    merged = Batch()
    for p in parts:
        merged.MergeFromString(p.SerializeToString())
    cols = ingest.decode_event_batch(b"".join(p.SerializeToString() for p in parts))
    _assert_columns_equal(cols, merged)


def test_writer_is_canonical():
    rng = np.random.default_rng(5)
    b = _random_batch(rng, 50)
    evs = [{"ts": (e.ts.seconds, e.ts.nanos) if e.HasField("ts") else None, "pid": e.pid, "tid": e.tid, "comm": e.comm, "syscall": e.syscall, "path": e.path,
            "new_path": e.new_path, "flags": e.flags, "ret_val": e.ret_val, "bytes": e.bytes} for e in b.events]
    _, Batch = classes()
    wire = ingest.encode_event_batch(evs)
    assert Batch.FromString(wire) == b
    assert wire == b.SerializeToString(deterministic=True)


def test_malformed_input_is_an_error_with_offset():
    b = _random_batch(np.random.default_rng(6), 4).SerializeToString()
    with pytest.raises(NerrfError, match="malformed EventBatch at byte offset"):
        ingest.decode_event_batch(b[:-3])
    with pytest.raises(NerrfError, match="malformed"):
        ingest.decode_event_batch(b"\x0a\xff\xff\xff\xff\xff\xff\xff\xff\xff\xff\xff")       # 11-byte varint
    with pytest.raises(NerrfError, match="malformed"):
        ingest.decode_event_batch(b"\x0b\x00")                                                  # group wire type
    assert ingest.decode_event_batch(b"").n == 0


@settings(max_examples=150, deadline=None)
@given(st.binary(max_size=200))
def test_fuzz_never_crashes_and_agrees_with_runtime_on_acceptance(data):
    _, Batch = classes()
    try:
        cols = ingest.decode_event_batch(data)
    except NerrfError:
        return
    # whatever we accept, the runtime accepts too and sees the same events (it may also accept inputs we reject,
    # e.g. groups; invalid UTF-8 is the one thing it rejects that we keep as bytes)
    try:
        b = Batch.FromString(data)
    except Exception:
        return
    assert cols.n == len(b.events)
    for i, e in enumerate(b.events):
        assert (cols.pid[i], cols.tid[i], cols.ret_val[i], cols.bytes[i]) == (e.pid, e.tid, e.ret_val, e.bytes)


def _graphs_equal(a, b):
    assert np.array_equal(a.rowptr, b.rowptr) and a.rowptr.dtype == b.rowptr.dtype
    assert np.array_equal(a.col, b.col) and np.array_equal(a.ew.view(np.uint32), b.ew.view(np.uint32))
    assert np.array_equal(a.x.view(np.uint32), b.x.view(np.uint32)), np.argwhere(a.x != b.x)[:5]
    assert a.meta["names"] == b.meta["names"]
    for k in ("node_kind", "label"):
        assert np.array_equal(a.meta[k], b.meta[k])
    assert np.array_equal(a.meta["size_mb"], b.meta["size_mb"])
    assert a.meta["t0"] == b.meta["t0"] and a.meta["span"] == b.meta["span"]


@pytest.mark.parametrize("merge", [True, False])
def test_columnar_constructor_equals_per_event_loader_on_lockbit_trace(merge):
    ev = trace_sim.lockbit_trace(n_files=25, seed=3, benign_files=10)
    for k, e in enumerate(ev):                           # give the rename-style events a target, and a second pid
        if e["event"] == "file_encrypt_complete":
            e["new_path"] = e["path"]
        e["pid"] = 454 if k % 7 else 455
    wire = ingest.encode_event_batch(ev)
    _, Batch = classes()
    ref_events = G.events_from_event_batch(Batch.FromString(wire))          # runtime parse -> per-event loader
    want = G.graph_from_events(ref_events, merge_renames=merge)
    cols = ingest.decode_event_batch(wire)
    got = ingest.graph_from_columns(cols, merge_renames=merge)
    _graphs_equal(got, want)
    assert ingest.events_from_columns(cols) == ref_events
    assert want.meta["label"].sum() == (25 if merge else 50)      # unmerged: x.dat and x.lockbit3 are two nodes


def test_columnar_constructor_on_unordered_random_stream():
    rng = np.random.default_rng(8)
    b = _random_batch(rng, 400)
    for e in b.events:                                   # keep times in one window and positive
        e.ts.seconds = 1_700_000_000 + int(rng.integers(0, 50)); e.ts.nanos = int(rng.integers(0, 4)) * 250_000_000
    wire = b.SerializeToString()
    for merge in (True, False):
        want = G.graph_from_events(G.events_from_event_batch(b), merge_renames=merge)
        got = ingest.graph_from_columns(ingest.decode_event_batch(wire), merge_renames=merge)
        _graphs_equal(got, want)


def test_replayed_m1_style_trace_scales_without_python_loops():
    ev = G.replicate_events(trace_sim.lockbit_trace(n_files=40, seed=1), 50)
    wire = ingest.encode_event_batch(ev)
    cols = ingest.decode_event_batch(wire)
    g = ingest.graph_from_columns(cols)
    assert cols.n == len(ev) and g.num_nodes > 50 * 40 and g.meta["label"].sum() == 50 * 40


def test_columnar_sequences_equal_per_event_sequences():
    from nerrf_b200 import pipeline
    ev = G.replicate_events(trace_sim.lockbit_trace(n_files=12, seed=5, benign_files=9), 3)
    rng = np.random.default_rng(0)
    hot = ev[40]["path"]                                  # one file with a long history: exercises the last-100 window
    t_hot = G._parse_ts(ev[-1]["timestamp"])
    for i in range(130):
        ev.append({"timestamp": t_hot + 0.01 * (i + 1) + float(rng.random()) * 0.001, "event": "write", "path": hot,
                   "size": int(rng.integers(1, 1 << 20)), "pid": 7})
    for e in ev:
        e.pop("phase", None)                              # the wire format carries no phase
        e["timestamp"] = G._parse_ts(e["timestamp"])
    cols = ingest.decode_event_batch(ingest.encode_event_batch(ev))
    events = ingest.events_from_columns(cols)
    g = ingest.graph_from_columns(cols)
    want_seq, want_len, want_nodes = pipeline.file_sequences(events, g)
    seq, lengths, nodes = ingest.sequences_from_columns(cols)
    assert np.array_equal(nodes, want_nodes) and np.array_equal(lengths, want_len) and lengths.max() == 100
    assert np.array_equal(seq.view(np.uint32), want_seq.view(np.uint32)), np.argwhere(seq != want_seq)[:5]


def _tracker_style_events():
    """What the real tracker emits (tracker/bpf/tracepoints.c:43-81): openat with a path, write WITHOUT a path, rename
    with path + new_path; two processes interleaved; one write before any open (it can name no file)."""
    ev, t = [], 1_700_000_000.0
    def emit(pid, syscall, path="", new_path="", size=0):
        nonlocal t
        t += 0.125
        ev.append({"timestamp": t, "event": syscall, "path": path, "new_path": new_path, "size": size, "pid": pid})
    emit(9, "write", size=77)                                        # no open yet: dropped
    for i in range(4):
        emit(7, "openat", f"/data/doc{i}.dat")
        emit(8, "openat", f"/srv/log{i}.txt")
        emit(7, "write", size=4096 + i)                              # -> /data/doc{i}.dat
        emit(8, "write", size=100 + i)                               # -> /srv/log{i}.txt
        emit(7, "write", size=4096)
        emit(7, "rename", f"/data/doc{i}.dat", f"/data/doc{i}.dat.lockbit3")
    emit(7, "openat", "/data/doc0.dat.lockbit3")                     # later event on the NEW name: same node
    return ev


@pytest.mark.parametrize("merge", [True, False])
def test_tracker_style_trace_pathless_writes_and_real_renames(merge):
    """path-less write events must not be interned as one global '' file node, and a real rename
    a.dat -> a.dat.lockbit3 must keep the identity and yield the rollback name."""
    ev = _tracker_style_events()
    wire = ingest.encode_event_batch(ev)
    cols = ingest.decode_event_batch(wire)
    assert cols.n == len(ev)
    want = G.graph_from_events(ingest.events_from_columns(cols), merge_renames=merge)
    got = ingest.graph_from_columns(cols, merge_renames=merge)
    _graphs_equal(got, want)
    names = got.meta["names"]
    assert "" not in names and "pid:9" not in names                 # the unattributable write is gone, with its process
    if merge:
        assert sorted(n for n in names if n.startswith("/data")) == [f"/data/doc{i}.dat.lockbit3" for i in range(4)]
        assert len(names) == 2 + 4 + 4                              # 2 pids, 4 docs (twin merged), 4 logs
        doc0 = names.index("/data/doc0.dat.lockbit3")
        assert got.x[doc0, 16] == 1.0                               # .lockbit bit from the rename target
        assert got.x[doc0, 5 + 5] == np.float32(np.log1p(2.0))      # both path-less writes were attributed to it
        assert got.x[doc0, 5 + 4] == np.float32(np.log1p(2.0))      # open of the old AND of the new name
        from nerrf_b200.ai.planner import emit
        assert emit.reversion_for(names[doc0]) == {"op": "rename", "from": "/data/doc0.dat.lockbit3", "to": "/data/doc0.dat.dat"}
    else:
        assert "/data/doc0.dat" in names and "/data/doc0.dat.lockbit3" in names
    # sequences: columnar == per-event, and every sequence belongs to a real file node
    from nerrf_b200 import pipeline
    w_seq, w_len, w_nodes = pipeline.file_sequences(ingest.events_from_columns(cols), want)
    seq, lengths, nodes = ingest.sequences_from_columns(cols, merge_renames=merge)
    assert np.array_equal(nodes, w_nodes) and np.array_equal(lengths, w_len)
    assert np.array_equal(seq.view(np.uint32), w_seq.view(np.uint32))
    assert all(got.meta["node_kind"][n] == 0 for n in nodes)


def test_observable_mode_hides_the_simulator_annotations():
    """features must not contain the ground-truth event kinds the label is defined by."""
    ev = trace_sim.lockbit_trace(n_files=8, seed=2, benign_files=6)
    g = G.graph_from_events(ev, observable=True)
    assert not g.x[:, 5:9].any()                                     # slots 0..3 (simulator-only kinds) are never counted
    assert g.meta["label"].sum() == 8                                # labels still come from the annotations
    gc = ingest.graph_from_columns(ingest.decode_event_batch(ingest.encode_event_batch(ev)), observable=True)
    assert np.array_equal(gc.x.view(np.uint32), g.x.view(np.uint32)) and np.array_equal(gc.meta["label"], g.meta["label"])
    from nerrf_b200 import pipeline
    seq, _, _ = pipeline.file_sequences(ev, g, observable=True)
    assert not seq[:, :, 0:4].any()
    seq2, _, _ = ingest.sequences_from_columns(ingest.decode_event_batch(ingest.encode_event_batch(ev)), observable=True)
    assert np.array_equal(seq2[:, :, :8], seq[:, :, :8])
