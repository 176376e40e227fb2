"""Streamed traces at fleet scale (BASELINE config 5: "End-to-end LockBit trace: streamed graph -> GNN anomaly score ->
MCTS rollback plan"): a columnar fleet-trace generator, the sliding window, and the tick loop around pipeline.run.

Reference anchors (prose + fixtures; the reference ships no graph constructor or planner):
  * "Sliding window (30-60 sec)", "Node merging (inode deduplication)"      docs/content/docs/architecture.mdx:39-41
  * pipeline order tracker -> graph -> AI models -> planner -> sandbox       docs/content/docs/architecture.mdx:12-86
  * one process's trace = the simulator's schema and phases                  benchmarks/m1/scripts/sim_lockbit_m1.py:24-36,
    benchmarks/m1/results/m1_trace.jsonl (149 events, 45 encrypted files)
  * SURVEY.md 8d cfg 5: "m1 trace replicated xK with pid/path renaming to a >= 1M-node stream"

Host code (numpy, no per-event Python): replicas are produced by tiling the columns of ONE decoded base trace and
patching a fixed-width placeholder in the path bytes."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import graph as G, ingest, trace_sim

_PLACEHOLDER = b"p0000000_"          # fixed width: 7 decimal digits patched per replica
_DIGITS = 7


def _base_columns(attacked: bool, n_files: int, benign_files: int, seed: int) -> ingest.EventColumns:
    ev = trace_sim.lockbit_trace(n_files=n_files if attacked else 0, seed=seed, benign_files=benign_files + (0 if attacked else n_files))
    if not attacked:                      # a benign process: no encryption phase, no ransom note
        ev = [e for e in ev if e["phase"] not in ("attack",) and e["event"] not in ("encryption_start", "encryption_complete")]
    for e in ev:
        slash = e["path"].rfind("/")
        if e["path"].startswith("/app/uploads/") and slash >= 0 and "." in e["path"][slash:]:
            e["path"] = e["path"][:slash + 1] + _PLACEHOLDER.decode() + e["path"][slash + 1:]   # per-process files; dirs stay shared
    return ingest.decode_event_batch(ingest.encode_event_batch(ev))


def _replicate(cols: ingest.EventColumns, ids: np.ndarray, t_shift: np.ndarray) -> ingest.EventColumns:
    """K copies of `cols`: replica j gets pid + ids[j], its placeholder digits = ids[j], its clock shifted by t_shift[j] s."""
    K, n = ids.shape[0], cols.n
    strings = {}
    for name, (off, data) in cols.strings.items():
        L = int(off[-1])
        d = np.tile(data, K).reshape(K, L) if L else np.zeros((K, 0), np.uint8)
        if L:
            raw = data.tobytes()
            pos, start = [], 0
            while True:
                k = raw.find(_PLACEHOLDER, start)
                if k < 0:
                    break
                pos.append(k + 1); start = k + 1
            if pos:
                pos = np.asarray(pos, np.int64)
                for dgt in range(_DIGITS):
                    digit = ((ids // 10 ** (_DIGITS - 1 - dgt)) % 10).astype(np.uint8) + np.uint8(ord("0"))
                    d[:, pos + dgt] = digit[:, None]
        new_off = np.empty(K * n + 1, np.int64)
        new_off[:-1] = (off[None, :-1] + (np.arange(K, dtype=np.int64) * L)[:, None]).ravel()
        new_off[-1] = K * L
        strings[name] = (new_off, d.reshape(-1))
    ts = cols.timestamp[None, :] + t_shift[:, None]
    sec = np.floor(ts)
    out = ingest.EventColumns(
        n=K * n, ts_sec=sec.astype(np.int64).ravel(), ts_nanos=np.round((ts - sec) * 1e9).astype(np.int32).clip(0, 999_999_999).ravel(),
        pid=(cols.pid[None, :].astype(np.int64) + ids[:, None]).astype(np.uint32).ravel(), tid=np.tile(cols.tid, K),
        flags=np.tile(cols.flags, K), ret_val=np.tile(cols.ret_val, K), bytes=np.tile(cols.bytes, K),
        event_slot=np.tile(cols.event_slot, K), path_flags=np.tile(cols.path_flags, K), strings=strings)
    return out


def concat_columns(parts) -> ingest.EventColumns:
    strings = {}
    for name in ingest.STRING_COLUMNS:
        offs, datas, base = [], [], 0
        for c in parts:
            off, data = c.strings[name]
            offs.append(off[:-1] + base); datas.append(data); base += int(off[-1])
        strings[name] = (np.concatenate(offs + [np.asarray([base], np.int64)]), np.concatenate(datas) if datas else np.zeros(0, np.uint8))
    cat = lambda k: np.concatenate([getattr(c, k) for c in parts])
    return ingest.EventColumns(n=sum(c.n for c in parts), strings=strings,
                               **{k: cat(k) for k in ("ts_sec", "ts_nanos", "pid", "tid", "flags", "ret_val", "bytes", "event_slot", "path_flags")})


def fleet_columns(n_procs: int, n_attacked: int, seed: int = 0, n_files: int = 45, benign_files: int = 50, jitter_s: float = 5.0,
                  return_pids: bool = False):
    """A fleet of `n_procs` processes, each with its own n_files + benign_files files under /app/uploads (the m1
    simulator's layout); `n_attacked` of them run the LockBit encryption phase on n_files of their files.
    -> (EventColumns, encrypted_paths: set of the .lockbit3 names a correct plan must rename back)."""
    rng = np.random.default_rng(seed)
    att = _base_columns(True, n_files, benign_files, seed)
    ben = _base_columns(False, n_files, benign_files, seed + 1)
    ids = rng.permutation(n_procs).astype(np.int64) + 1000
    shift = rng.random(n_procs) * jitter_s
    parts = []
    if n_attacked:
        parts.append(_replicate(att, ids[:n_attacked], shift[:n_attacked]))
    if n_procs > n_attacked:
        parts.append(_replicate(ben, ids[n_attacked:], shift[n_attacked:]))
    cols = concat_columns(parts)
    enc = set()
    if n_attacked:
        a = parts[0]
        off, data = a.strings["path"]
        idx = np.nonzero((a.path_flags & 8) != 0)[0]
        raw = data.tobytes()
        enc = {raw[off[i]:off[i + 1]].decode() for i in idx.tolist()}
    if return_pids:                       # the ransomware processes, by the node name the graph constructor gives them
        return cols, enc, {"pid:%d" % p_ for p_ in np.unique(parts[0].pid).tolist()} if n_attacked else set()
    return cols, enc


def window(cols: ingest.EventColumns, t_lo: float, t_hi: float) -> ingest.EventColumns:
    """The events with t_lo < t <= t_hi (the sliding window of architecture.mdx:39-41), stored order preserved."""
    ts = cols.timestamp
    keep = np.nonzero((ts > t_lo) & (ts <= t_hi))[0]
    strings = {name: ingest._gather_strings(*cols.strings[name], keep) for name in ingest.STRING_COLUMNS}
    sc = {k: getattr(cols, k)[keep] for k in ("ts_sec", "ts_nanos", "pid", "tid", "flags", "ret_val", "bytes", "event_slot", "path_flags")}
    return ingest.EventColumns(n=int(keep.shape[0]), strings=strings, **sc)


# ---------------------------------------------------------------------------------------------- device-resident stream
_M64 = (1 << 64) - 1


def _mix64(h: int) -> int:
    h ^= h >> 33; h = (h * 0xff51afd7ed558ccd) & _M64; h ^= h >> 33; h = (h * 0xc4ceb9fe1a85ec53) & _M64; h ^= h >> 33
    return h or 1


def name_hash(name) -> int:
    """The 64-bit name hash of csrc/intern_device.cu (nerrf_trace_name_hash) on the host, as a SIGNED int64: FNV-1a over
    the bytes seeded with the length, murmur3 finaliser; "pid:<n>" names hash their pid."""
    if isinstance(name, str) and name.startswith("pid:") and name[4:].isdigit():
        h = _mix64((int(name[4:]) + 0x9E3779B97F4A7C15) & _M64)
    else:
        b = name.encode("utf-8") if isinstance(name, str) else bytes(name)
        h = (0xcbf29ce484222325 ^ len(b)) & _M64
        for c in b:
            h = ((h ^ c) * 0x100000001b3) & _M64
        h = _mix64(h)
    return h - (1 << 64) if h >= (1 << 63) else h


class LazyNames:
    """Node names of a window graph, decoded on demand (a window holds ~10^6 nodes, a plan names a few thousand)."""

    def __init__(self, cols: ingest.EventColumns, name_event: np.ndarray, name_which: np.ndarray):
        self.cols, self.name_event, self.name_which = cols, name_event, name_which

    def __len__(self):
        return int(self.name_event.shape[0])

    def __getitem__(self, v):
        e, w = int(self.name_event[v]), int(self.name_which[v])
        return "pid:%d" % int(self.cols.pid[e]) if w == 2 else self.cols.text("path" if w == 0 else "new_path", e)


class DeviceStream:
    """The event stream resident in HBM ("window, inode/path dedup via hash, sort-by-dst -> CSR,
    feature counts" on the GPU; docs/content/docs/architecture.mdx:39-41).  Columns are uploaded ONCE; a sliding window
    is a slice of the time-sorted index array, and everything per tick -- node interning (hash table), per-node features,
    edge assembly, CSR -- runs on the device.  The host keeps the strings only to NAME the few nodes a plan touches."""

    def __init__(self, cols: ingest.EventColumns, device="cuda", observable=True, merge_renames=True):
        import torch
        self.torch = torch
        cols = ingest.resolve_columns(cols)                   # path-less (write) events follow the pid's open file
        self.cols, self.observable, self.merge = cols, observable, bool(merge_renames)
        self.dev = dev = torch.device(device)
        ts = cols.timestamp
        self.order = np.argsort(ts, kind="stable")
        self.ts_sorted = ts[self.order]
        up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dt)).to(dev)
        n = cols.n
        self.n = n
        (poff, pdata), (goff, gdata) = cols.strings["path"], cols.strings["new_path"]
        self.d_order = up(self.order, np.int64)
        self.d_ts = up(ts, np.float64)
        self.d_pid = up(cols.pid.astype(np.uint32).view(np.int32), np.int32)
        self.d_poff, self.d_goff = up(poff, np.int64), up(goff, np.int64)
        self.d_pdata = up(pdata if pdata.size else np.zeros(1, np.uint8), np.uint8)
        self.d_gdata = up(gdata if gdata.size else np.zeros(1, np.uint8), np.uint8)
        raw = cols.event_slot
        self.d_raw_slot = up(raw, np.uint8)
        self.d_slot = up(np.asarray(G.OBSERVABLE_SLOT, np.uint8)[raw] if observable else raw, np.uint8)
        self.d_bytes = up(cols.bytes.astype(np.int64), np.int64)
        pf = cols.path_flags
        if self.merge:                                        # the renamed twin is the same node: its .lockbit bit counts
            pf = pf | (ingest.path_flags_of(goff, gdata) & np.uint8(1))
        self.d_pf = up(pf, np.uint8)
        self.d_pf_raw = up(cols.path_flags, np.uint8)         # the sequences use the event's own path flags
        self.cap = max((2 if self.merge else 3) * n, 1)
        self.node_p = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        self.node_f = torch.empty_like(self.node_p); self.node_g = torch.empty_like(self.node_p)
        self.kind = torch.zeros(self.cap, dtype=torch.int8, device=dev)
        self.name_event = torch.zeros(self.cap, dtype=torch.int64, device=dev)
        self.name_which = torch.zeros(self.cap, dtype=torch.int8, device=dev)
        import ctypes as C
        from . import _lib
        need = C.c_int64()
        _lib.check(_lib.lib().nerrf_trace_intern_device_workspace_bytes(max(n, 1), self.cap, C.byref(need)), "intern_device_workspace_bytes")
        self._ws = torch.empty(need.value + 256, dtype=torch.uint8, device=dev)
        self._ws_bytes = need.value

    def span(self):
        if self.n == 0:
            return 0.0, 0.0
        return float(self.ts_sorted[0]), float(self.ts_sorted[-1])

    def window_graph(self, t_lo: float, t_hi: float, window_s: float):
        """The temporal graph of the events with t_lo < t <= t_hi (device tensors), or None when there are none.
        Same construction as ingest.graph_from_columns(stream.window(cols, t_lo, t_hi), device=...)."""
        import ctypes as C
        from . import _lib
        torch = self.torch
        lo = int(np.searchsorted(self.ts_sorted, t_lo, "right")); hi = int(np.searchsorted(self.ts_sorted, t_hi, "right"))
        nw = hi - lo
        if nw <= 0:
            return None
        sel = self.d_order[lo:hi]                                   # the window = a slice of the time-sorted index array
        nn = C.c_int64()
        ws_ptr = (self._ws.data_ptr() + 255) & ~255
        with torch.cuda.device(self.dev):
            _lib.check(_lib.lib().nerrf_trace_intern_device(
                nw, self.n, _lib.ptr(sel), _lib.ptr(self.d_pid), _lib.ptr(self.d_poff), _lib.ptr(self.d_pdata), _lib.ptr(self.d_goff),
                _lib.ptr(self.d_gdata), int(self.merge), _lib.ptr(self.node_p), _lib.ptr(self.node_f), _lib.ptr(self.node_g),
                C.byref(nn), _lib.ptr(self.kind), _lib.ptr(self.name_event), _lib.ptr(self.name_which), self.cap,
                C.c_void_p(ws_ptr), self._ws_bytes, _lib.current_stream_ptr()), "nerrf_trace_intern_device")
        N = int(nn.value)
        t0 = float(self.ts_sorted[lo]); span = max(float(self.ts_sorted[hi - 1]) - t0, 1e-6)
        window = window_s or max(span, G.WINDOW)
        P, F = self.node_p[sel], self.node_f[sel]                   # by rank: the window's events in time order
        t = self.d_ts[sel] - t0
        kind_d = self.kind[:N].contiguous()
        Gn = None
        if not self.merge:
            Gn = self.node_g[sel]
            if not bool((Gn >= 0).any()):
                Gn = None
        x, label, size_mb = G.node_features_device(P, F, Gn, t, self.d_slot[sel], self.d_bytes[sel], self.d_pf[sel], kind_d, window)
        tt = t.to(torch.float32)
        if Gn is None:                                              # edges in the host loader's per-event order: p->f, f->p
            src = torch.stack([P, F], 1).reshape(-1); dst = torch.stack([F, P], 1).reshape(-1)
            te = tt.repeat_interleave(2)
        else:                                                       # [, f->g, g->f]
            keep = torch.stack([torch.ones_like(P, dtype=torch.bool)] * 2 + [Gn >= 0] * 2, 1).reshape(-1)
            src = torch.stack([P, F, F, Gn], 1).reshape(-1)[keep]; dst = torch.stack([F, P, Gn, F], 1).reshape(-1)[keep]
            te = tt.repeat_interleave(4)[keep]
        rowptr, col, ew, perm = G.build_csr_device(src.contiguous(), dst.contiguous(), te.contiguous(), torch.ones_like(te), N,
                                                   t_ref=float(span), tau=G.TAU, return_perm=True)
        lab = label
        if self.observable:                                         # the kernel saw the folded slots: labels come from the annotations
            raw = self.d_raw_slot[sel]
            lab = (torch.bincount(F[(raw == 1) | (raw == 2)].long(), minlength=N) > 0)
        nh = torch.empty(N, dtype=torch.int64, device=self.dev)
        with torch.cuda.device(self.dev):
            _lib.check(_lib.lib().nerrf_trace_name_hash(N, _lib.ptr(self.name_event), _lib.ptr(self.name_which), _lib.ptr(self.d_pid),
                                                        _lib.ptr(self.d_poff), _lib.ptr(self.d_pdata), _lib.ptr(self.d_goff),
                                                        _lib.ptr(self.d_gdata), _lib.ptr(nh), _lib.current_stream_ptr()),
                       "nerrf_trace_name_hash")
        meta = {"kind": "trace", "names": LazyNames(self.cols, self.name_event[:N].cpu().numpy(), self.name_which[:N].cpu().numpy()),
                "node_kind": kind_d.cpu().numpy().astype(np.int64), "t0": t0, "span": span, "merge_renames": self.merge,
                "label": lab.cpu().numpy().astype(np.int64), "size_mb": size_mb.cpu().numpy(), "device": str(self.dev),
                "name_hash": nh, "events": nw, "window_ranks": (lo, hi), "perm": perm if Gn is None else None}
        return G.TemporalGraph(rowptr, col, ew, x, meta)

    def sequences_device(self, g: G.TemporalGraph, only_nodes):
        """LSTM sequences of the given file nodes of window graph `g`, built ON THE DEVICE (nerrf_trace_sequences) from the
        resident columns and the graph's own CSR rows: -> (seq fp32 [A, T_MAX, 16], lengths int32 [A]) CUDA tensors in the
        order of `only_nodes`.  Needs the two-edges-per-event graph of merge mode (g.meta['perm'])."""
        import ctypes as C
        from . import _lib
        from .ai.models import lstm
        torch = self.torch
        perm = g.meta.get("perm")
        if perm is None:
            raise _lib.NerrfError("sequences_device needs a merge-mode window graph (two edges per event)")
        lo, hi = g.meta["window_ranks"]
        cand = torch.as_tensor(np.asarray(only_nodes, np.int64)).to(self.dev)
        A = int(cand.shape[0])
        seq = torch.empty(A, lstm.T_MAX, lstm.D_IN, dtype=torch.float32, device=self.dev)
        ln = torch.empty(A, dtype=torch.int32, device=self.dev)
        with torch.cuda.device(self.dev):
            _lib.check(_lib.lib().nerrf_trace_sequences(_lib.ptr(cand), A, _lib.ptr(g.rowptr), int(g.rowptr.dtype == torch.int64),
                                                        _lib.ptr(perm), _lib.ptr(self.d_order[lo:hi]), _lib.ptr(self.d_ts),
                                                        _lib.ptr(self.d_slot), _lib.ptr(self.d_bytes), _lib.ptr(self.d_pf_raw),
                                                        float(g.meta["t0"]), float(g.meta["span"]), lstm.T_MAX, _lib.ptr(seq),
                                                        _lib.ptr(ln), _lib.current_stream_ptr()), "nerrf_trace_sequences")
        return seq, ln

    def sequences(self, g: G.TemporalGraph, only_nodes):
        """LSTM sequences of the given file nodes of window graph `g` (ingest.sequences_core on the window's events)."""
        lo, hi = g.meta["window_ranks"]
        order = self.order[lo:hi]
        if "_F_host" not in g.meta:
            g.meta["_F_host"] = self.node_f[self.d_order[lo:hi]].cpu().numpy().astype(np.int64)
        return ingest.sequences_core(self.cols, order, g.meta["_F_host"], self.ts_sorted[lo:hi], g.meta["t0"], g.meta["span"],
                                     g.num_nodes, None, self.observable, only_nodes)


@dataclass
class TickResult:
    t_hi: float
    events: int
    nodes: int
    edges: int
    planned: list                  # paths this tick's plan reverts (new ones only)
    timings_ms: dict
    truncated: bool = False


@dataclass
class StreamingPlanner:
    """Tick loop: every `tick_s` seconds of trace time the events of the last `window_s` seconds become a temporal graph
    (features + CSR on the GPU), GraphSAGE_T scores it, the top-A anomalous files that were not reverted yet go
    through the LSTM and the MCTS planner, and the resulting reversions are added to the running undo plan."""
    model: object
    scorer: object
    window_s: float = 60.0
    tick_s: float = 30.0
    top_a: int = 4096
    n_rollouts: int = 1024
    depth: int = 32
    iterations: int = 8
    commit_per_search: int = 64
    kill_candidates: bool = False  # also propose "kill process" actions (planner spec v1: reversions need their writer gone)
    device: str = "cuda"
    dist_ctx: object = None        # pipeline.DistContext for the multi-GPU form
    reverted: set = field(default_factory=set)
    _reverted_hash: set = field(default_factory=set)
    ingest_ms: float = 0.0         # one-time: stream columns -> HBM (DeviceStream)
    killed: set = field(default_factory=set)
    ticks: list = field(default_factory=list)

    def run(self, cols: ingest.EventColumns):
        """cols: the event stream (multi-GPU: only rank 0's copy is read -- rank 0 does the host-side ingest of every tick
        and broadcasts the device graph, the exclusion mask and the candidates' sequences; names stay on rank 0)."""
        from . import pipeline
        import time
        import torch
        ctx = self.dist_ctx
        multi = ctx is not None and ctx.world > 1
        lead = not multi or ctx.rank == 0
        dev = torch.device(self.device)
        if multi:
            import torch.distributed as dist

            def bcast(t, shape=None, dtype=None):
                """broadcast a tensor from rank 0 (shape / dtype known to everybody, or sent first)"""
                if shape is None:
                    hdr = torch.tensor(list(t.shape) + [-1] * (4 - t.dim()) if lead else [0] * 4, device=dev, dtype=torch.int64)
                    dist.broadcast(hdr, 0)
                    shape = [int(v) for v in hdr.tolist() if v >= 0]
                buf = t.to(dev).contiguous() if lead else torch.empty(shape, device=dev, dtype=dtype)
                dist.broadcast(buf, 0)
                return buf
        ds = None
        if lead:
            # the stream becomes device resident ONCE (in production: batch by batch as the tracker delivers it); from here
            # on a tick touches no string on the host except the names of the nodes its plan reverts
            a = time.perf_counter()
            ds = DeviceStream(cols, self.device, observable=True)
            torch.cuda.synchronize()
            self.ingest_ms = (time.perf_counter() - a) * 1e3
            span = torch.tensor(list(ds.span()), dtype=torch.float64, device=dev)
        else:
            span = torch.zeros(2, dtype=torch.float64, device=dev)
        if multi:
            dist.broadcast(span, 0)
        t0, t1 = float(span[0]), float(span[1])
        t_hi = t0
        while t_hi < t1:
            t_hi = min(t_hi + self.tick_s, t1)
            tm = {}
            names, g, n_events = None, None, 0
            a = time.perf_counter()
            if lead:
                # window = a slice of the time-sorted index array; interning (hash table), features, edges, CSR on the GPU
                g = ds.window_graph(t_hi - self.window_s, t_hi, self.window_s)
                torch.cuda.synchronize()
                n_events = g.meta["events"] if g is not None else 0
            if multi:
                ne = torch.tensor([n_events], device=dev); dist.broadcast(ne, 0); n_events = int(ne)
            if n_events == 0:
                continue
            if lead:
                names = g.meta["names"]
            if multi:                                      # the device graph travels over NVLink; the strings do not
                x = bcast(g.x if lead else None, dtype=torch.float32)
                rp = bcast(g.rowptr.to(torch.int64) if lead else None, dtype=torch.int64).to(torch.int32)
                col = bcast(g.col if lead else None, dtype=torch.int32)
                ew = bcast(g.ew if lead else None, dtype=torch.float32)
                kind_t = bcast(torch.from_numpy(np.asarray(g.meta["node_kind"]).astype(np.int64)) if lead else None, dtype=torch.int64)
                size_t = bcast(torch.from_numpy(np.asarray(g.meta["size_mb"], np.float32)) if lead else None, dtype=torch.float32)
                if not lead:
                    g = G.TemporalGraph(rp, col, ew, x, {"kind": "trace", "node_kind": kind_t.cpu().numpy(), "size_mb": size_t.cpu().numpy()})
            tm["graph_build"] = (time.perf_counter() - a) * 1e3
            kind = np.asarray(g.meta["node_kind"])
            nodes = np.nonzero(kind == 0)[0]

            # LSTM sequences are built lazily, for the top-A candidates only (a window holds ~10^6 file nodes)
            def seq(cand, g=g):
                if lead and not multi:                      # built on the device from the resident columns, in candidate order
                    sq_d, ln_d = ds.sequences_device(g, cand)
                    return sq_d, ln_d, None
                if lead:
                    sq, ln, have = ds.sequences(g, cand)
                if multi:
                    sq_t = bcast(torch.from_numpy(sq) if lead else None, dtype=torch.float32)
                    ln_t = bcast(torch.from_numpy(ln) if lead else None, dtype=torch.int32)
                    hv_t = bcast(torch.from_numpy(have) if lead else None, dtype=torch.int64)
                    return sq_t.cpu().numpy(), ln_t.cpu().numpy(), hv_t.cpu().numpy()
                return sq, ln, have

            new, truncated = [], False
            for rnd in range(4):
                # a planning pass proposes at most 32 process kills (planner spec v1: guards live in state word 0); a tick
                # with more suspicious processes than that plans again over what is still unreverted
                skip = None
                if lead and self.reverted:               # files already reverted by an earlier tick: matched by name hash on the device
                    rh = torch.tensor(sorted(self._reverted_hash), dtype=torch.int64, device=dev)
                    skip = torch.isin(g.meta["name_hash"][torch.from_numpy(nodes).to(dev)], rh).cpu().numpy()
                if multi:
                    flag = torch.tensor([1 if (lead and skip is not None) else 0], device=dev); dist.broadcast(flag, 0)
                    if int(flag):
                        skip = bcast(torch.from_numpy(skip) if lead else None, shape=[nodes.shape[0]], dtype=torch.bool).cpu().numpy()
                res = pipeline.run(g, seq, None, nodes, self.model, self.scorer, top_a=self.top_a, n_rollouts=self.n_rollouts,
                                   depth=self.depth, iterations=self.iterations, device=self.device, exclude=skip,
                                   commit_per_search=self.commit_per_search, dist_ctx=self.dist_ctx,
                                   kill_candidates=self.kill_candidates)
                for k, v in res.timings_ms.items():
                    tm[k] = tm.get(k, 0.0) + float(v)
                got = []
                if lead:
                    self.killed.update(names[n] for n in res.plan_nodes if kind[n] == 1)
                    got = [names[n] for n in res.plan_nodes if kind[n] == 0]
                    got = [nm for nm in dict.fromkeys(got) if nm not in self.reverted]
                    self.reverted.update(got)
                    self._reverted_hash.update(name_hash(nm) for nm in got)
                    new += got
                more = bool(self.kill_candidates and res.n_kill >= 32 and got)
                if multi:
                    flag = torch.tensor([int(more)], device=dev); dist.broadcast(flag, 0); more = bool(int(flag))
                truncated = truncated or bool(res.plan.truncated)
                if not more:
                    break
            self.ticks.append(TickResult(t_hi - t0, n_events, g.num_nodes, g.num_edges, new, tm, truncated))
        return self.ticks
