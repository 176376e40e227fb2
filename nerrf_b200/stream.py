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
        if lead:
            ts = cols.timestamp
            span = torch.tensor([float(ts.min()), float(ts.max())], dtype=torch.float64, device=dev)
        else:
            span = torch.zeros(2, dtype=torch.float64, device=dev)
        if multi:
            dist.broadcast(span, 0)
        t0, t1 = float(span[0]), float(span[1])
        t_hi = t0
        while t_hi < t1:
            t_hi = min(t_hi + self.tick_s, t1)
            tm = {}
            names, w, n_events = None, None, 0
            if lead:
                a = time.perf_counter()
                w = window(cols, t_hi - self.window_s, t_hi)
                tm["window"] = (time.perf_counter() - a) * 1e3
                n_events = w.n
            if multi:
                ne = torch.tensor([n_events], device=dev); dist.broadcast(ne, 0); n_events = int(ne)
            if n_events == 0:
                continue
            a = time.perf_counter()
            if lead:
                g = ingest.graph_from_columns(w, device=self.device, observable=True, window=self.window_s)
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
            def seq(cand, w=w):
                if lead:
                    sq, ln, have = ingest.sequences_from_columns(w, observable=True, only_nodes=cand)
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
                if lead and self.reverted:
                    skip = np.asarray([names[n] in self.reverted for n in nodes.tolist()], bool)
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
                    got = [names[n] for n in res.plan_nodes if kind[n] == 0 and names[n] not in self.reverted]
                    self.reverted.update(got)
                    new += got
                more = bool(self.kill_candidates and res.n_kill >= 32 and got)
                if multi:
                    flag = torch.tensor([int(more)], device=dev); dist.broadcast(flag, 0); more = bool(int(flag))
                truncated = truncated or bool(res.plan.truncated)
                if not more:
                    break
            self.ticks.append(TickResult(t_hi - t0, n_events, g.num_nodes, g.num_edges, new, tm, truncated))
        return self.ticks
