"""End-to-end detection -> planning pipeline on one GPU (BASELINE config 5 in miniature / at scale).

    events (EventBatch / TRACE json) --graph.py--> temporal graph --GraphSAGE_T--> node anomaly scores
      --> top-A candidate file nodes --lstm--> encrypt_probability --rewards.Actions--> mcts.plan --> undo plan

Reference anchors: the intended stack of docs/content/docs/architecture.mdx:12-86 (graph constructor ->
AI models -> planner -> sandbox) and threat-model.mdx:141-223 (phases 2-5).  Everything on the hot path runs
through the C-ABI kernels; the host side here is glue + timing.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field

import numpy as np
import torch

from . import graph as G
from .ai.models import GraphSAGE_T, lstm
from .ai.planner import mcts
from .ai.planner.rewards import Actions, KIND_KILL as RW_KIND_KILL

_SEQ_EVENT_SLOT = {"file_created": 0, "file_encrypt_start": 1, "file_encrypt_complete": 2, "ransom_note_created": 3,
                   "openat": 4, "write": 5, "rename": 6}


def file_sequences(events, g: G.TemporalGraph, t_max=lstm.T_MAX, observable=False):
    """Per-file event sequences for the LSTM (the last `t_max` events of each file node, oldest first).
    Features (D_in = 16): one-hot event kind (8), log1p(size)/20, dt to the previous event of the file (s, clipped),
    time since trace start / window, .lockbit flag, /tmp|/proc flag, 3 spare (the simulator's `phase` annotation is
    ground truth, not an observable: it is deliberately NOT a feature, so wire-format traces give the same sequences)."""
    merge = g.meta.get("merge_renames", True)
    keys = g.meta.get("file_keys")
    if keys is None:                                       # graphs from the columnar constructor: rebuild the key map from the names
        keys = {(G._stem(n) if merge else n): i for i, n in enumerate(g.meta["names"])}
    key_of = (lambda p_: G._stem(p_)) if merge else (lambda p_: p_)
    evs = G.resolve_event_paths(events)                    # path-less (write) events follow the pid's open file
    t0 = G._parse_ts(evs[0]["timestamp"]) if evs else 0.0
    span = max(G._parse_ts(evs[-1]["timestamp"]) - t0, 1e-6) if evs else 1.0
    per_file: dict = {}
    for e in evs:
        node = keys.get(key_of(e["path"]))
        if node is None or g.meta["node_kind"][node] != 0:
            continue
        per_file.setdefault(node, []).append(e)
    nodes = sorted(per_file)
    seq = np.zeros((len(nodes), t_max, lstm.D_IN), np.float32)
    lengths = np.zeros(len(nodes), np.int32)
    for i, n in enumerate(nodes):
        es = per_file[n][-t_max:]
        lengths[i] = len(es)
        prev = None
        for k, e in enumerate(es):
            t = G._parse_ts(e["timestamp"])
            slot = _SEQ_EVENT_SLOT.get(e["event"], 7)
            seq[i, k, G.OBSERVABLE_SLOT[slot] if observable else slot] = 1.0
            seq[i, k, 8] = np.log1p(float(e.get("size", 0) or 0)) / 20.0
            seq[i, k, 9] = min(t - prev, 10.0) if prev is not None else 0.0
            seq[i, k, 10] = (t - t0) / span
            seq[i, k, 11] = 1.0 if ".lockbit" in e["path"] else 0.0
            seq[i, k, 12] = 1.0 if e["path"].startswith(("/tmp", "/proc")) else 0.0
            prev = t
    return seq, lengths, np.asarray(nodes, np.int64)


@dataclass
class PipelineResult:
    node_score: torch.Tensor
    candidates: np.ndarray            # node ids of the A candidate files
    probs: torch.Tensor               # [A, 2] encrypt_probability, ransomware_score
    plan: mcts.Plan
    plan_nodes: list                  # node ids to revert, in plan order
    timings_ms: dict = field(default_factory=dict)
    n_kill: int = 0                   # the first n_kill candidates are process nodes ("kill process" actions)


@dataclass
class DistContext:
    """Multi-GPU form of the pipeline (SURVEY.md 8e rows LSTM / MCTS / e2e): one process per GPU, every rank holds the
    graph; GraphSAGE-T runs 1-D edge-block sharded with the fused per-layer exchange, the own-row scores are
    all-gathered, the LSTM batch is split across ranks and all-gathered, MCTS is root-parallel (rank r searches with
    seed + r, root statistics summed in rank order) -- every rank ends with the same plan."""
    rank: int
    world: int
    exchange: str = "broadcast"      # NCCL all-gather-v per layer: window graphs change size every tick, so the peer-mapped
                                     # buffers of the fused exchange (sized by N, collectively allocated) would be rebuilt per tick


def _sage_scores(g_dev, model, ctx):
    x, rp, col, ew = g_dev
    if ctx is None or ctx.world == 1:
        return model(x, rp, col, ew)[1]
    import torch.distributed as dist
    from . import dist as ND
    ss = ND.ShardedSage(model, rp, col, ew, ctx.rank, ctx.world, x.device, exchange=ctx.exchange)
    ss.set_x(x)
    score = ss.step()
    cuts = ss.shard.cuts
    parts = [torch.empty(int(cuts[p + 1] - cuts[p]), device=x.device) for p in range(ctx.world)]
    dist.all_gather(parts, score[int(cuts[ctx.rank]):int(cuts[ctx.rank + 1])].contiguous())
    return torch.cat(parts)


def _lstm_probs(seq_model, seq_sel, len_sel, dev, ctx):
    if torch.is_tensor(seq_sel):                                       # built on the device (stream.DeviceStream.sequences_device)
        return seq_model(seq_sel.to(dev), len_sel.to(dev))
    if ctx is None or ctx.world == 1 or seq_sel.shape[0] < ctx.world:
        return seq_model(torch.from_numpy(seq_sel).to(dev), torch.from_numpy(len_sel).to(dev))
    import torch.distributed as dist
    a = seq_sel.shape[0]
    per = (a + ctx.world - 1) // ctx.world
    lo, hi = min(ctx.rank * per, a), min((ctx.rank + 1) * per, a)
    mine = torch.zeros(per, 2, device=dev)
    if hi > lo:
        mine[:hi - lo] = seq_model(torch.from_numpy(seq_sel[lo:hi]).to(dev), torch.from_numpy(len_sel[lo:hi]).to(dev))
    parts = [torch.empty(per, 2, device=dev) for _ in range(ctx.world)]
    dist.all_gather(parts, mine)
    return torch.cat(parts)[:a]


def process_kill_candidates(g: G.TemporalGraph, candidates: np.ndarray, max_kills: int = 32):
    """Non-revert undo candidates derived from the graph (threat-model.mdx:208-222 "Kill process", cost 10): the
    process nodes that wrote the candidate files, the ones with the most candidate files first.
    -> (pid_nodes [k], guard [A] = index into pid_nodes of the process that wrote each candidate file, or -1)."""
    kind = torch.as_tensor(np.asarray(g.meta["node_kind"]))
    rp = (g.rowptr.cpu() if torch.is_tensor(g.rowptr) else torch.from_numpy(np.asarray(g.rowptr))).long()
    col = (g.col.cpu() if torch.is_tensor(g.col) else torch.from_numpy(np.asarray(g.col))).long()
    cand = torch.from_numpy(np.asarray(candidates, np.int64))
    # the most recent in-edge of a file node (rows are time-sorted) whose source is a process = its last writer
    owner = np.full(cand.shape[0], -1, np.int64)
    last = rp[cand + 1] - 1
    has = (rp[cand + 1] > rp[cand]).numpy()
    src = col[last.clamp_min(0)]
    is_proc = (kind[src] == 1).numpy() & has
    owner[is_proc] = src.numpy()[is_proc]
    for i in np.nonzero(has & ~is_proc)[0].tolist():          # rare: the last toucher is a file (rename link): scan back
        n = int(cand[i]); s_ = col[rp[n]:rp[n + 1]]; s_ = s_[kind[s_] == 1]
        if s_.numel():
            owner[i] = int(s_[-1])
    pids, counts = np.unique(owner[owner >= 0], return_counts=True)
    if pids.size == 0:
        return np.zeros(0, np.int64), np.full(cand.shape[0], -1, np.int64)
    chosen = pids[np.lexsort((pids, -counts))[:max_kills]]
    slot = np.full(int(max(pids.max(), 0)) + 1, -1, np.int64); slot[chosen] = np.arange(chosen.shape[0])
    guard = np.where(owner >= 0, slot[np.maximum(owner, 0)], -1)
    return chosen, guard


def run(g: G.TemporalGraph, seq, lengths, seq_nodes, model: GraphSAGE_T, seq_model, top_a=1024, confidence=None,
        plan_steps=None, n_rollouts=4096, depth=50, iterations=16, device="cuda", exclude=None, commit_per_search=1,
        dist_ctx: DistContext | None = None, kill_candidates: bool = False) -> PipelineResult:
    """seq/lengths/seq_nodes: LSTM inputs for (a superset of) the candidate file nodes (file_sequences()).
    seq may instead be a CALLABLE nodes -> (seq, lengths, nodes_with_sequences): the sequences are then built only for
    the top-A candidates (large windows: a million file nodes, a few thousand candidates); seq_nodes = the file nodes.
    confidence: optional override of p_a per NODE (e.g. ground-truth labels when no trained weights exist).
    exclude: bool [len(seq_nodes)] -- sequences whose file was already reverted by an earlier tick of the stream.
    kill_candidates: also propose "kill process" actions (cost 10) for the processes that wrote the candidate files;
    a file reversion only sticks once its writer is gone (planner spec v1)."""
    dev = torch.device(device)
    tm = {}
    sync = torch.cuda.synchronize

    t0 = time.perf_counter()
    # numpy arrays (host constructor) are uploaded; CUDA tensors (ingest.graph_from_columns(device=...)) are used as is
    x, rp, col, ew = ((a if torch.is_tensor(a) else torch.from_numpy(a)).to(dev, non_blocking=True)
                      for a in (g.x, g.rowptr, g.col, g.ew))
    sync(); tm["h2d_graph"] = (time.perf_counter() - t0) * 1e3

    t0 = time.perf_counter()
    score = _sage_scores((x, rp, col, ew), model, dist_ctx)
    sync(); tm["graphsage_t"] = (time.perf_counter() - t0) * 1e3

    t0 = time.perf_counter()
    seq_nodes_t = torch.from_numpy(np.asarray(seq_nodes)).to(dev)
    cand_score = score[seq_nodes_t]
    if exclude is not None and np.asarray(exclude).any():
        cand_score = cand_score.masked_fill(torch.from_numpy(np.asarray(exclude)).to(dev), -1.0)
    a = min(top_a, int(seq_nodes_t.numel() if exclude is None else (~np.asarray(exclude)).sum()))
    if a == 0:
        return PipelineResult(score, np.zeros(0, np.int64), torch.zeros(0, 2, device=dev), mcts.Plan([], [0.0], []), [], dict(tm, total=sum(tm.values())))
    cand_local = torch.topk(cand_score, a).indices                   # most anomalous files that have sequences
    candidates = seq_nodes_t[cand_local]
    sync(); tm["top_a"] = (time.perf_counter() - t0) * 1e3

    t0 = time.perf_counter()
    if callable(seq):
        want = candidates.cpu().numpy()
        sq, ln, have = seq(want)                                       # sequences of the candidates only
        if have is None:                                               # already in candidate order (device builder: CUDA tensors)
            seq_sel, len_sel = sq, ln
        else:                                                          # sorted by node id (host builder)
            pos = np.searchsorted(have, want)
            assert have.shape[0] == want.shape[0] and np.array_equal(have[pos], want), "every candidate file has events"
            seq_sel, len_sel = sq[pos], ln[pos]
        tm["sequences"] = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
    else:
        cl = cand_local.cpu().numpy()
        seq_sel, len_sel = seq[cl], lengths[cl]
    probs = _lstm_probs(seq_model, seq_sel, len_sel, dev, dist_ctx)
    sync(); tm["lstm"] = (time.perf_counter() - t0) * 1e3

    t0 = time.perf_counter()
    cand = candidates.cpu().numpy()
    size_mb = np.maximum(g.meta.get("size_mb", np.ones(g.num_nodes, np.float32))[cand], 0.01)
    n_kill = 0
    if confidence is not None:
        p = np.asarray(confidence, np.float32)[cand]
        actions = Actions(p, size_mb, np.ones(a, np.float32))
    elif kill_candidates:
        pids, guard = process_kill_candidates(g, cand)
        n_kill = int(pids.shape[0])
        a_rev = min(a, 4096 - n_kill)                                  # kill actions occupy the first slots (spec v1: guards < 32)
        p_rev = 0.5 * (score[candidates[:a_rev]].cpu().numpy() + probs[:a_rev, 0].cpu().numpy())
        # confidence that a process is malicious = mean confidence of the candidate files it wrote (the model's node head
        # is trained on file nodes only)
        gr = guard[:a_rev]
        p_kill = np.array([p_rev[gr == k].mean() if (gr == k).any() else 0.0 for k in range(n_kill)], np.float32)
        p_all = np.concatenate([p_kill, p_rev]).astype(np.float32)
        size_all = np.concatenate([np.zeros(n_kill, np.float32), size_mb[:a_rev]])
        kind = np.concatenate([np.full(n_kill, RW_KIND_KILL, np.int64), np.zeros(a_rev, np.int64)])
        cost = np.where(kind == RW_KIND_KILL, 10.0, 1.0).astype(np.float32)
        actions = Actions(p_all, size_all, cost, kind, guard=np.concatenate([np.full(n_kill, -1, np.int64), guard[:a_rev]]))
        cand = np.concatenate([pids, cand[:a_rev]])
    else:
        actions = Actions.from_scores(score[candidates], probs[:, 0], size_mb, np.zeros(a, np.int64))
    merge = None
    if dist_ctx is not None and dist_ctx.world > 1:
        from . import dist as ND
        merge = lambda n, w: ND.root_parallel_search(lambda off: (n, w), dist_ctx.rank, dist_ctx.world)
    pl = mcts.plan(actions, max_steps=plan_steps, n_rollouts=n_rollouts, depth=depth, seed=dist_ctx.rank if dist_ctx else 0,
                   iterations=iterations, device=dev, commit_per_search=commit_per_search, merge=merge)
    sync(); tm["mcts_plan"] = (time.perf_counter() - t0) * 1e3
    tm["total"] = sum(tm.values())
    res = PipelineResult(score, cand, probs, pl, [int(cand[i]) for i in pl.actions], tm)
    res.n_kill = n_kill
    return res
