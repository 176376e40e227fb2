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
from .ai.planner.rewards import Actions

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


def run(g: G.TemporalGraph, seq, lengths, seq_nodes, model: GraphSAGE_T, seq_model, top_a=1024, confidence=None,
        plan_steps=None, n_rollouts=4096, depth=50, iterations=16, device="cuda") -> PipelineResult:
    """seq/lengths/seq_nodes: LSTM inputs for (a superset of) the candidate file nodes (file_sequences()).
    confidence: optional override of p_a per NODE (e.g. ground-truth labels when no trained weights exist)."""
    dev = torch.device(device)
    tm = {}
    sync = torch.cuda.synchronize

    t0 = time.perf_counter()
    # numpy arrays (host constructor) are uploaded; CUDA tensors (ingest.graph_from_columns(device=...)) are used as is
    x, rp, col, ew = ((a if torch.is_tensor(a) else torch.from_numpy(a)).to(dev, non_blocking=True)
                      for a in (g.x, g.rowptr, g.col, g.ew))
    sync(); tm["h2d_graph"] = (time.perf_counter() - t0) * 1e3

    t0 = time.perf_counter()
    h, score = model(x, rp, col, ew)
    sync(); tm["graphsage_t"] = (time.perf_counter() - t0) * 1e3

    t0 = time.perf_counter()
    seq_nodes_t = torch.from_numpy(np.asarray(seq_nodes)).to(dev)
    a = min(top_a, seq_nodes_t.numel())
    cand_local = torch.topk(score[seq_nodes_t], a).indices           # most anomalous files that have sequences
    candidates = seq_nodes_t[cand_local]
    sync(); tm["top_a"] = (time.perf_counter() - t0) * 1e3

    t0 = time.perf_counter()
    cl = cand_local.cpu().numpy()
    probs = seq_model(torch.from_numpy(seq[cl]).to(dev), torch.from_numpy(lengths[cl]).to(dev))
    sync(); tm["lstm"] = (time.perf_counter() - t0) * 1e3

    t0 = time.perf_counter()
    cand = candidates.cpu().numpy()
    size_mb = np.maximum(g.meta.get("size_mb", np.ones(g.num_nodes, np.float32))[cand], 0.01)
    if confidence is not None:
        p = np.asarray(confidence, np.float32)[cand]
        actions = Actions(p, size_mb, np.ones(a, np.float32))
    else:
        actions = Actions.from_scores(score[candidates], probs[:, 0], size_mb, np.zeros(a, np.int64))
    pl = mcts.plan(actions, max_steps=plan_steps, n_rollouts=n_rollouts, depth=depth,
                   iterations=iterations, device=dev)
    sync(); tm["mcts_plan"] = (time.perf_counter() - t0) * 1e3
    tm["total"] = sum(tm.values())
    return PipelineResult(score, cand, probs, pl, [int(cand[i]) for i in pl.actions], tm)
