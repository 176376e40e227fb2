"""`nerrf undo` for the hot path: trace in, undo plan out (reference CLI `nerrf undo --id <attack>`,
ROADMAP.md:28,86 -- scheduled, never written).

    python -m nerrf_b200.undo --trace events.pb   --weights weights.pt --out plan.json --shell plan.sh
    python -m nerrf_b200.undo --trace trace.jsonl --train-epochs 25    --out plan.json

trace: serialized nerrf.trace.EventBatch bytes (*.pb; proto/trace.proto:47-49, e.g. a buffered StreamEvents stream) or
the simulator's TRACE json lines (*.jsonl; benchmarks/m1/scripts/sim_lockbit_m1.py:24-36).  Stages: ingest -> temporal
graph (features + CSR on the GPU for *.pb) -> GraphSAGE_T.forward -> top-A candidates -> lstm.forward -> mcts plan ->
emit.  Everything between the graph and the plan runs in the CUDA kernels; there is no CPU fallback.
"""
from __future__ import annotations

import argparse
import sys
import time

import torch

from . import graph as G, ingest, pipeline
from .ai import train as T
from .ai.models import GraphSAGE_T
from .ai.models.lstm import LSTMScorer
from .ai.planner import emit


def load_trace(path: str, device="cuda"):
    """-> (graph, seq, lengths, seq_nodes, n_events)"""
    if path.endswith(".jsonl") or path.endswith(".json"):
        events = G.read_trace_jsonl(path)
        g = G.graph_from_events(events, observable=True)          # the features ai/train.py trains on
        seq, lengths, nodes = pipeline.file_sequences(events, g, observable=True)
        return g, seq, lengths, nodes, len(events)
    with open(path, "rb") as f:
        cols = ingest.decode_event_batch(f.read())
    g = ingest.graph_from_columns(cols, device=device, observable=True)
    seq, lengths, nodes = ingest.sequences_from_columns(cols, observable=True)
    return g, seq, lengths, nodes, cols.n


def load_models(weights: str | None, train_epochs: int, layers: int = 2, log=print):
    model, scorer = GraphSAGE_T(G.F_IN, 128, layers), LSTMScorer()
    if weights:
        ck = torch.load(weights, map_location="cpu")
        model = GraphSAGE_T(G.F_IN, 128, int(ck.get("layers", layers)))
        model.load_state_dict(ck["sage"])
        if ck.get("lstm") is not None:
            scorer.load_state_dict(ck["lstm"])
    elif train_epochs > 0:
        log(f"no --weights: training {train_epochs} epochs on the simulated toy set (ai/train.py)")
        torch.manual_seed(0)
        T.train(model, scorer, T.toy_set(range(100, 104)), epochs=train_epochs)
    else:
        raise SystemExit("no weights: pass --weights FILE (from `python -m nerrf_b200.ai.train --out FILE`) or --train-epochs N")
    return model, scorer


def main(argv=None) -> dict:
    ap = argparse.ArgumentParser(prog="python -m nerrf_b200.undo", description=__doc__.split("\n")[0])
    ap.add_argument("--trace", required=True); ap.add_argument("--id", default=None, help="attack id carried into the plan")
    ap.add_argument("--weights", default=None); ap.add_argument("--train-epochs", type=int, default=0)
    ap.add_argument("--out", default=None, help="plan JSON (default: stdout)"); ap.add_argument("--shell", default=None)
    ap.add_argument("--top-a", type=int, default=1024); ap.add_argument("--rollouts", type=int, default=1024)
    ap.add_argument("--depth", type=int, default=50); ap.add_argument("--iterations", type=int, default=16)
    ap.add_argument("--max-steps", type=int, default=None,
                    help="cap on the number of reversions (default: every candidate may be reverted); a capped plan that "
                         "leaves improving candidates is marked truncated and is not auto-approvable")
    ap.add_argument("--device", default="cuda")
    a = ap.parse_args(argv)
    if not torch.cuda.is_available():
        raise SystemExit("nerrf_b200.undo needs a CUDA device (no CPU fallback for the hot path)")
    log = lambda *m: print(*m, file=sys.stderr)
    t0 = time.perf_counter()
    g, seq, lengths, nodes, n_events = load_trace(a.trace, a.device)
    t_ingest = (time.perf_counter() - t0) * 1e3
    model, scorer = load_models(a.weights, a.train_epochs, log=log)
    model.to(a.device); scorer.to(a.device)
    res = pipeline.run(g, seq, lengths, nodes, model, scorer, top_a=min(a.top_a, 4096), n_rollouts=a.rollouts, depth=a.depth,
                       iterations=a.iterations, plan_steps=a.max_steps, device=a.device)
    plan = emit.from_pipeline(g, res, attack_id=a.id)
    plan["stats"] = {"events": int(n_events), "nodes": int(g.num_nodes), "edges": int(g.num_edges),
                     "candidates": int(len(res.candidates)), "ingest_ms": t_ingest,
                     "timings_ms": {k: float(v) for k, v in res.timings_ms.items()}}
    text = emit.to_json(plan)
    if a.out:
        with open(a.out, "w") as f:
            f.write(text + "\n")
    else:
        print(text)
    if a.shell:
        with open(a.shell, "w") as f:
            f.write(emit.to_shell(plan))
    log(f"{n_events} events -> {g.num_nodes} nodes / {g.num_edges} edges -> {len(plan['steps'])} reversions "
        f"(ingest {t_ingest:.1f} ms, hot path {res.timings_ms['total']:.1f} ms)")
    return plan


if __name__ == "__main__":
    main()
