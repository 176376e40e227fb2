"""Undo-plan emitter: the step AFTER the hot path -- host code only.

Turns a planner result (ordered action indices over candidate file nodes) into the artefact the
reference hands to its sandbox / `nerrf undo` CLI: an ordered list of file reversions with the
validation rule the sandbox applies.

Reference anchors (prose + the M1 harness; no emitter exists in the reference):
  * planner output "Undo plan (file reversions, process kills)"  docs/content/docs/architecture.mdx:62-72
  * sandbox workflow "apply undo operations ... validate: md5sum matches pre-attack version ...
    approve if all checks pass"                                    docs/content/docs/architecture.mdx:75-86
  * what a reversion IS for the LockBit simulator: `x.lockbit3` -> `x.dat`
                                                                   benchmarks/m1/scripts/m1_rollback.sh:95-112
  * CLI `nerrf undo --id <attack>`                                 ROADMAP.md:28,86
[ours]: the JSON field names (the reference defines none); version field so they can change.
"""
from __future__ import annotations

import json
import shlex

PLAN_VERSION = 1
ENCRYPTED_SUFFIX = ".lockbit3"      # benchmarks/m1/scripts/m1_rollback.sh:95 (`for f in *.lockbit3`)
RESTORED_SUFFIX = ".dat"            # benchmarks/m1/scripts/m1_rollback.sh:98 (`mv "$f" "${base}.dat"`)


def _has_control_chars(path: str) -> bool:
    return any(ord(ch) < 0x20 or ord(ch) == 0x7F for ch in path)


def reversion_for(path: str) -> dict:
    """One file reversion.  An encrypted twin is renamed back (the M1 harness' recovery); anything else
    has no in-place inverse and is restored from the pre-attack snapshot the sandbox cloned.
    Paths come from attacker-controlled file names in the trace: an empty path is refused (it names no file;
    the tracker zeroes `path` on write events, tracker/bpf/tracepoints.c:62-64), and a path with control
    characters is flagged so that no consumer ever pastes it into a line-oriented format."""
    if not path:
        raise ValueError("cannot emit a reversion for an empty path")
    if path.endswith(ENCRYPTED_SUFFIX):
        r = {"op": "rename", "from": path, "to": path[: -len(ENCRYPTED_SUFFIX)] + RESTORED_SUFFIX}
    else:
        r = {"op": "restore_snapshot", "path": path}
    if _has_control_chars(path):
        r["hostile_name"] = True       # exact bytes kept (JSON escapes them); to_shell() only ever passes them as a quoted argument
    return r


def undo_plan(names, plan_nodes, scores=None, attack_id=None, size_mb=None, node_score=None, truncated=False,
              remaining_candidates=0) -> dict:
    """names: node id -> path (TemporalGraph.meta['names']); plan_nodes: node ids to revert, in plan order
    (PipelineResult.plan_nodes); scores: Plan.scores (exact reward after each commit, scores[0] = before).
    truncated / remaining_candidates: the planner stopped at its step limit while improving candidates remained --
    such a plan must not be auto-approved as complete."""
    steps = []
    for k, n in enumerate(plan_nodes):
        n = int(n)
        step = {"order": k, "node": n, **reversion_for(names[n]),
                "validate": {"check": "md5", "against": "pre_attack_snapshot"}}
        if size_mb is not None:
            step["size_mb"] = float(size_mb[n])
        if node_score is not None:
            step["anomaly_score"] = float(node_score[n])
        if scores is not None and k + 1 < len(scores):
            step["reward_after"] = float(scores[k + 1])
        steps.append(step)
    out = {"version": PLAN_VERSION, "kind": "nerrf.undo_plan", "attack_id": attack_id, "steps": steps,
           "approve_if": "all_checks_pass" if not truncated else "all_checks_pass_and_operator_confirms_partial_plan",
           "truncated": bool(truncated), "remaining_candidates": int(remaining_candidates)}
    if scores:
        out["reward_before"] = float(scores[0])
        out["reward_after"] = float(scores[min(len(plan_nodes), len(scores) - 1)])
    return out


def to_json(plan: dict) -> str:
    return json.dumps(plan, indent=2, sort_keys=True)


def to_shell(plan: dict) -> str:
    """The same plan as the commands the M1 rollback script runs (m1_rollback.sh:95-99), for dry runs."""
    lines = ["#!/bin/sh", "set -e"]
    for s in plan["steps"]:
        if s["op"] == "rename":
            lines.append("mv -- %s %s" % (shlex.quote(s["from"]), shlex.quote(s["to"])))
        else:
            # never a comment: a newline inside a path would end the comment and start a command.  `:` is the
            # POSIX no-op; the path is only ever a single-quoted ARGUMENT (newlines inside '...' are literal).
            lines.append(": restore-from-pre-attack-snapshot %s" % shlex.quote(s["path"]))
    return "\n".join(lines) + "\n"


def from_pipeline(g, result, attack_id=None) -> dict:
    """g: TemporalGraph built from a trace; result: pipeline.PipelineResult."""
    score = result.node_score.detach().cpu().numpy() if hasattr(result.node_score, "detach") else result.node_score
    return undo_plan(g.meta["names"], result.plan_nodes, scores=result.plan.scores, attack_id=attack_id,
                     size_mb=g.meta.get("size_mb"), node_score=score, truncated=getattr(result.plan, "truncated", False),
                     remaining_candidates=getattr(result.plan, "remaining_candidates", 0))
