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
    exchange: str = "p2p"


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


def process_kill_candidates(g: G.TemporalGraph, candidates: np.ndarray, score, max_kills: int = 32):
    """Non-revert undo candidates derived from the graph (threat-model.mdx:208-222 "Kill process", cost 10): the
    process nodes that touched the candidate files, most anomalous first.  -> (pid_nodes [k], guard [A] = index into
    pid_nodes of the process that wrote each candidate file, or -1)."""
    kind = np.asarray(g.meta["node_kind"])
    rp = g.rowptr.cpu().numpy() if torch.is_tensor(g.rowptr) else np.asarray(g.rowptr)
    col = g.col.cpu().numpy() if torch.is_tensor(g.col) else np.asarray(g.col)
    owner = np.full(candidates.shape[0], -1, np.int64)
    for i, n in enumerate(candidates.tolist()):              # in-edges of a file node come from the processes that touched it
        src = col[rp[n]:rp[n + 1]]
        src = src[kind[src] == 1]
        if src.size:
            owner[i] = int(src[-1])                          # the most recent writer
    pids, counts = np.unique(owner[owner >= 0], return_counts=True)
    if pids.size == 0:
        return np.zeros(0, np.int64), np.full(candidates.shape[0], -1, np.int64)
    sc = score[torch.from_numpy(pids).to(score.device)].cpu().numpy() if torch.is_tensor(score) else np.asarray(score)[pids]
    order = np.lexsort((pids, -counts, -sc))[:max_kills]
    chosen = pids[order]
    slot = {int(p_): k for k, p_ in enumerate(chosen.tolist())}
    guard = np.asarray([slot.get(int(o), -1) for o in owner.tolist()], np.int64)
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
        sq, ln, have = seq(want)                                       # sequences of the candidates only (sorted by node id)
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
        pids, guard = process_kill_candidates(g, cand, score)
        n_kill = int(pids.shape[0])
        a_rev = min(a, 4096 - n_kill)                                  # kill actions occupy the first slots (spec v1: guards < 32)
        p_rev = 0.5 * (score[candidates[:a_rev]].cpu().numpy() + probs[:a_rev, 0].cpu().numpy())
        p_all = np.concatenate([score[torch.from_numpy(pids).to(dev)].cpu().numpy(), p_rev]).astype(np.float32)
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
