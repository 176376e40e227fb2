"""Undo-plan emitter (host code): plan nodes -> ordered reversions, JSON, shell."""
import json
import os
import subprocess

import numpy as np

from nerrf_b200 import graph as G, trace_sim
from nerrf_b200.ai.planner import emit


def test_reversion_matches_m1_rollback_semantics():
    # benchmarks/m1/scripts/m1_rollback.sh:95-99: base="${f%.lockbit3}"; mv "$f" "${base}.dat"
    r = emit.reversion_for("/app/uploads/m1_file_07.lockbit3")
    assert r == {"op": "rename", "from": "/app/uploads/m1_file_07.lockbit3", "to": "/app/uploads/m1_file_07.dat"}
    assert emit.reversion_for("/app/uploads/report.dat") == {"op": "restore_snapshot", "path": "/app/uploads/report.dat"}


def test_plan_over_simulated_trace_round_trips_through_json_and_shell(tmp_path):
    events = trace_sim.lockbit_trace(n_files=6, seed=1)
    g = G.graph_from_events(events)
    names = g.meta["names"]
    enc = [n for n, nm in enumerate(names) if nm.endswith(".lockbit3")]
    assert len(enc) == 6
    scores = [-10.0] + [-10.0 + 1.5 * (k + 1) for k in range(len(enc))]
    plan = emit.undo_plan(names, enc, scores=scores, attack_id="m1-sim", size_mb=g.meta["size_mb"])
    assert plan["version"] == 1 and plan["attack_id"] == "m1-sim" and len(plan["steps"]) == 6
    assert [s["order"] for s in plan["steps"]] == list(range(6))
    assert plan["reward_before"] == -10.0 and plan["reward_after"] == scores[-1]
    assert all(s["validate"]["check"] == "md5" for s in plan["steps"])
    assert json.loads(emit.to_json(plan)) == plan
    # dry run of the shell form in a scratch directory: every encrypted twin comes back as .dat
    for s in plan["steps"]:
        rel = s["from"].lstrip("/")
        os.makedirs(tmp_path / os.path.dirname(rel), exist_ok=True)
        (tmp_path / rel).write_text("x")
    script = emit.to_shell(plan).replace(" /", " ./").replace("'/", "'./")
    subprocess.run(["sh", "-c", script], cwd=tmp_path, check=True)
    for s in plan["steps"]:
        assert (tmp_path / s["to"].lstrip("/")).exists() and not (tmp_path / s["from"].lstrip("/")).exists()


def test_shell_quotes_hostile_names():
    plan = emit.undo_plan({0: "/data/a b; rm -rf $HOME.lockbit3"}, [0])
    line = emit.to_shell(plan).splitlines()[2]
    assert line == "mv -- '/data/a b; rm -rf $HOME.lockbit3' '/data/a b; rm -rf $HOME.dat'"


def test_partial_scores_and_numpy_ids():
    plan = emit.undo_plan(["/f0.lockbit3", "/f1"], np.asarray([1, 0]), scores=[-3.0, -2.0])
    assert plan["steps"][0]["op"] == "restore_snapshot" and plan["steps"][0]["reward_after"] == -2.0
    assert "reward_after" not in plan["steps"][1] and plan["reward_after"] == -2.0


def test_hostile_names_cannot_inject_commands(tmp_path):
    """a newline in a path used to end the `# restore ...` comment line and start a command."""
    marker = tmp_path / "PWNED"
    names = {0: f"/data/a\ntouch {marker} #.dat", 1: f"/data/b'; touch {marker}; '.dat", 2: f"/data/$(touch {marker}).dat",
             3: f"/data/`touch {marker}`\r.dat"}
    plan = emit.undo_plan(names, [0, 1, 2, 3])
    assert all(s["op"] == "restore_snapshot" for s in plan["steps"])
    assert plan["steps"][0]["hostile_name"] and plan["steps"][3]["hostile_name"] and "hostile_name" not in plan["steps"][1]
    script = emit.to_shell(plan)
    assert not any(l.lstrip().startswith("#") for l in script.splitlines()[1:])      # no comment lines carry data
    subprocess.run(["sh", "-c", script], cwd=tmp_path, check=True)
    assert not marker.exists()
    # a rename of a hostile encrypted name: the file is moved, nothing else happens
    enc = "we ird\n$(touch PWNED)'name.lockbit3"                                   # cwd = tmp_path: same marker file
    (tmp_path / enc).write_text("x")
    plan = emit.undo_plan({0: enc}, [0])
    subprocess.run(["sh", "-c", emit.to_shell(plan)], cwd=tmp_path, check=True)
    assert (tmp_path / (enc[:-len(".lockbit3")] + ".dat")).exists() and not marker.exists()
    assert json.loads(emit.to_json(plan))["steps"][0]["from"] == enc                 # exact bytes survive the JSON form


def test_empty_path_is_refused_and_truncation_is_reported():
    import pytest
    with pytest.raises(ValueError):
        emit.reversion_for("")
    plan = emit.undo_plan(["/a.lockbit3"], [0], truncated=True, remaining_candidates=7)
    assert plan["truncated"] and plan["remaining_candidates"] == 7 and plan["approve_if"] != "all_checks_pass"
    assert emit.undo_plan(["/a.lockbit3"], [0])["approve_if"] == "all_checks_pass"
