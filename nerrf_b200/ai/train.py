"""ai.train -- joint GraphSAGE-T + BiLSTM training.

Reference surface: `ai/train.py` "joint GNN+LSTM training script" (README.md:75 -- named, never written);
gate "GNN / LSTM ROC-AUC >= 0.90 on the toy set" (ROADMAP.md:26,62-69); labels / classes
docs/content/docs/threat-model.mdx:176-203,277-292.

Training is NOT the north-star hot path (that is inference: GraphSAGE_T.forward, lstm.forward, mcts.search,
rewards.score -- CUDA only).  With `--device cuda` the GraphSAGE-T half trains through the library's own kernels: the
fused tcgen05 layer forward and the hand-written backward (`nerrf_sage_layer_bwd`, csrc/sage_bwd.cu) behind a
torch.autograd.Function (nerrf_b200.ai.autograd); on the CPU it differentiates a plain-torch statement of the same frozen
spec (also the thing the CUDA gradients are tested against).  The BiLSTM half is torch.nn.LSTM under autograd on either
device.  Both update the very nn.Parameters the CUDA inference kernels read (GraphSAGE_T.weights / biases / node_w /
node_b; LSTMScorer.lstm / head).  Inference never calls anything in this file, and nothing here is a fallback for it:
GraphSAGE_T.forward / LSTMScorer.forward still refuse CPU tensors.

    python -m nerrf_b200.ai.train --traces 8 --epochs 200 --out weights.pt
"""
from __future__ import annotations

import argparse

import numpy as np
import torch
from torch import nn
from torch.nn.utils.rnn import pack_padded_sequence

from .. import graph as G, pipeline, trace_sim
from .models import GraphSAGE_T
from .models.lstm import LSTMScorer


# ------------------------------------------------------------------ differentiable restatements (autograd only)
def sage_node_logits(model: GraphSAGE_T, x, rowptr, col, ew):
    """Node-head logits of GraphSAGE-T with autograd: h' = ReLU([h || m] W + b), m = weighted mean over in-edges,
    logit = h . w_n + b_n  (the spec in ai/models/graphsage_t.py; sigmoid(logit) is what the CUDA forward returns)."""
    N = x.shape[0]
    deg = (rowptr[1:] - rowptr[:-1]).long()
    dst = torch.repeat_interleave(torch.arange(N, device=x.device), deg)
    src = col.long()
    wsum = torch.zeros(N, device=x.device, dtype=x.dtype).index_add_(0, dst, ew).clamp_min(1e-12)
    h = x
    for W, b in zip(model.weights, model.biases):
        agg = torch.zeros(N, h.shape[1], device=x.device, dtype=x.dtype).index_add_(0, dst, h[src] * ew[:, None])
        h = torch.relu(torch.cat([h, agg / wsum[:, None]], 1) @ W + b)
    return h @ model.node_w + model.node_b


def node_logits(model: GraphSAGE_T, ex):
    """Node-head logits of one example with gradients: on a CUDA device through the library's own kernels (fused tcgen05
    forward + csrc/sage_bwd.cu backward, nerrf_b200.ai.autograd), on the CPU through the plain-torch restatement above."""
    if ex["x"].is_cuda:
        from . import autograd as AG
        tg = ex.get("_train_graph")
        if tg is None:
            tg = ex["_train_graph"] = AG.TrainGraph(ex["rowptr"], ex["col"], ex["ew"])
        return AG.sage_node_logits(model, ex["x"], tg)
    return sage_node_logits(model, ex["x"], ex["rowptr"], ex["col"], ex["ew"])


def lstm_logits(scorer: LSTMScorer, seq, lengths):
    """Head logits [B, 2] of the BiLSTM with autograd (torch.nn.LSTM on packed sequences = the frozen spec)."""
    packed = pack_padded_sequence(seq, lengths.cpu().to(torch.int64), batch_first=True, enforce_sorted=False)
    _, (hn, _) = scorer.lstm(packed)
    return scorer.head(torch.cat([hn[-2], hn[-1]], 1))


def roc_auc(scores, labels) -> float:
    """Area under the ROC curve by the rank statistic (ties get the average rank)."""
    s = np.asarray(scores, np.float64); y = np.asarray(labels).astype(bool)
    n1, n0 = int(y.sum()), int((~y).sum())
    if n1 == 0 or n0 == 0:
        return float("nan")
    order = np.argsort(s, kind="stable")
    ranks = np.empty(s.size, np.float64)
    sorted_s = s[order]
    i = 0
    while i < s.size:
        j = i
        while j + 1 < s.size and sorted_s[j + 1] == sorted_s[i]:
            j += 1
        ranks[order[i:j + 1]] = 0.5 * (i + j) + 1.0
        i = j + 1
    return float((ranks[y].sum() - n1 * (n1 + 1) / 2.0) / (n1 * n0))


# ------------------------------------------------------------------ toy set (simulator-schema LockBit traces)
def make_example(seed: int, n_files: int = 30, benign_files: int = 40, window=None) -> dict:
    """One labelled example: a simulated LockBit trace -> graph tensors + per-file sequences.  Labels: a file node
    is positive iff it was encrypted (graph.graph_from_events meta['label'], from the simulator's annotations).
    Features are OBSERVABLE-ONLY (graph.OBSERVABLE_SLOT): the annotation event kinds that define the label
    (file_encrypt_start / _complete, ...) are folded onto the syscalls the tracker can actually see
    (openat / write / rename), so the gate below is not satisfied by reading the label back from the input.
    What remains is what a wire trace carries: write / rename counts, byte counts, timing, the .lockbit extension
    ("extension pattern", threat-model.mdx:178-184).  The toy set is still easy; see DESIGN.md 2.7."""
    ev = trace_sim.lockbit_trace(n_files=n_files, seed=seed, benign_files=benign_files)
    if window is not None:
        # what the streaming pipeline sees (nerrf_b200.stream): only the events of one sliding window, so a file's
        # history may start with its encryption and the clock features are relative to the window
        t_all = [G._parse_ts(e["timestamp"]) for e in ev]
        lo, hi = t_all[0] + window[0], t_all[0] + window[1]
        ev = [e for e, t_ in zip(ev, t_all) if lo < t_ <= hi]
    g = G.graph_from_events(ev, observable=True, window=None if window is None else float(window[1] - window[0]))
    seq, lengths, nodes = pipeline.file_sequences(ev, g, observable=True)
    t = torch.from_numpy
    return {"x": t(g.x), "rowptr": t(g.rowptr), "col": t(g.col), "ew": t(g.ew),
            "label": t(g.meta["label"].astype(np.float32)), "is_file": t(g.meta["node_kind"] == 0),
            "seq": t(seq), "lengths": t(lengths), "seq_label": t(g.meta["label"][nodes].astype(np.float32)),
            "graph": g, "events": ev, "seq_nodes": nodes}


def toy_set(seeds, windows: int = 2, **kw):
    """Per seed: the whole trace plus `windows` random 60 s sliding windows of it (the shape the streamed pipeline feeds
    the models, docs/content/docs/architecture.mdx:39-41)."""
    rng = np.random.default_rng(1234)
    out = []
    for s in seeds:
        nf, nb = int(rng.integers(10, 40)), int(rng.integers(10, 60))
        out.append(make_example(int(s), n_files=nf, benign_files=nb, **kw))
        span = 3.0 + 0.3 * (nf + nb) + 2.0 + 1.41 * nf                # recon + seeding + encryption phases (trace_sim)
        for _ in range(windows):
            hi = float(rng.uniform(0.45, 1.05) * span)
            ex = make_example(int(s), n_files=nf, benign_files=nb, window=(hi - 60.0, hi), **kw)
            if ex["seq"].shape[0] > 1 and 0 < float(ex["label"].sum()):
                out.append(ex)
    return out


def _to(ex, device):
    return {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in ex.items()}


def train(model: GraphSAGE_T, scorer: LSTMScorer | None, examples, epochs: int = 150, lr: float = 3e-3, device="cpu",
          log=None):
    """Joint loss = BCE(node logits over file nodes) + BCE(LSTM encrypt_probability logit); Adam; full batch per trace."""
    device = torch.device(device)
    model.to(device)
    params = list(model.parameters())
    if scorer is not None:
        scorer.to(device)
        params += list(scorer.parameters())
    opt = torch.optim.Adam(params, lr=lr)
    data = [_to(e, device) for e in examples]
    bce = nn.functional.binary_cross_entropy_with_logits
    for ep in range(epochs):
        total = 0.0
        for ex in data:
            opt.zero_grad()
            logit = node_logits(model, ex)
            m = ex["is_file"]
            pos = ex["label"][m].sum().clamp_min(1.0)
            loss = bce(logit[m], ex["label"][m], pos_weight=((m.sum() - pos) / pos).clamp(0.2, 5.0))
            if scorer is not None and ex["seq"].shape[0]:
                sl = lstm_logits(scorer, ex["seq"], ex["lengths"])
                loss = loss + bce(sl[:, 0], ex["seq_label"]) + 0.5 * bce(sl[:, 1], ex["seq_label"])
            loss.backward()
            opt.step()
            total += float(loss.detach())
        if log and (ep % 25 == 0 or ep == epochs - 1):
            log(f"epoch {ep:4d}  loss {total / len(data):.4f}")
    return model, scorer


@torch.no_grad()
def evaluate(model: GraphSAGE_T, scorer: LSTMScorer | None, examples, device="cpu") -> dict:
    """ROC-AUC of the node scores over file nodes and of the LSTM encrypt_probability over sequences (autograd
    restatement; tests/test_gpu_train.py checks the CUDA inference path gives the same scores)."""
    device = torch.device(device)
    ns, nl, ss, sl_ = [], [], [], []
    for ex in (_to(e, device) for e in examples):
        logit = sage_node_logits(model.to(device), ex["x"], ex["rowptr"], ex["col"], ex["ew"])
        m = ex["is_file"]
        ns.append(torch.sigmoid(logit[m]).cpu().numpy()); nl.append(ex["label"][m].cpu().numpy())
        if scorer is not None and ex["seq"].shape[0]:
            p = torch.sigmoid(lstm_logits(scorer.to(device), ex["seq"], ex["lengths"]))
            ss.append(p[:, 0].cpu().numpy()); sl_.append(ex["seq_label"].cpu().numpy())
    out = {"gnn_auc": roc_auc(np.concatenate(ns), np.concatenate(nl))}
    if ss:
        out["lstm_auc"] = roc_auc(np.concatenate(ss), np.concatenate(sl_))
    return out


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--traces", type=int, default=8); ap.add_argument("--epochs", type=int, default=150)
    ap.add_argument("--layers", type=int, default=2); ap.add_argument("--lr", type=float, default=3e-3)
    ap.add_argument("--device", default="cpu"); ap.add_argument("--out", default=None)
    ap.add_argument("--no-lstm", action="store_true")
    a = ap.parse_args(argv)
    train_set = toy_set(range(100, 100 + a.traces)); test_set = toy_set(range(900, 903))
    model = GraphSAGE_T(G.F_IN, 128, a.layers); scorer = None if a.no_lstm else LSTMScorer()
    print("before:", evaluate(model, scorer, test_set, a.device))
    train(model, scorer, train_set, a.epochs, a.lr, a.device, log=print)
    print("held-out:", evaluate(model, scorer, test_set, a.device))
    if a.out:
        torch.save({"sage": model.state_dict(), "lstm": None if scorer is None else scorer.state_dict(),
                    "layers": a.layers}, a.out)
        print("wrote", a.out)


if __name__ == "__main__":
    main()
