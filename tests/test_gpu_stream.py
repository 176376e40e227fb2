"""BASELINE config 5 in miniature (needs a B200): a fleet trace streamed tick by tick through the sliding window, the
GPU graph constructor, GraphSAGE_T, the LSTM on the top-A candidates and the MCTS planner with process-kill candidates
(planner spec v1) -- the running undo plan must end up reverting the encrypted files and killing their writers."""
import numpy as np
import pytest
import torch

from nerrf_b200 import stream
from nerrf_b200.ai import train as T
from nerrf_b200.ai.models import GraphSAGE_T
from nerrf_b200.ai.models.lstm import LSTMScorer

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def models():
    torch.manual_seed(0)
    model, scorer = GraphSAGE_T(32, 128, 2), LSTMScorer()
    T.train(model, scorer, T.toy_set(range(100, 104)), epochs=25, lr=3e-3)
    return model.cuda(), scorer.cuda()


@pytest.mark.parametrize("kills", [False, True])
def test_streamed_fleet_trace_plans_the_reversions(models, kills):
    model, scorer = models
    cols, enc, bad = stream.fleet_columns(120, 5, seed=9, return_pids=True)
    sp = stream.StreamingPlanner(model, scorer, top_a=1024, n_rollouts=512, depth=24, iterations=8, commit_per_search=32,
                                 kill_candidates=kills)
    ticks = sp.run(cols)
    assert len(ticks) >= 3 and ticks[0].planned == []            # nothing is encrypted in the first 30 s
    planned = set(sp.reverted)
    tp = len(planned & enc)
    assert tp / len(enc) >= 0.95, (tp, len(enc))                  # recall: (nearly) every encrypted file is renamed back
    assert tp / max(len(planned), 1) >= 0.90                      # precision: the ransom note is the usual extra
    assert all(t.timings_ms["graphsage_t"] > 0 and "sequences" in t.timings_ms for t in ticks if t.planned)
    if kills:
        assert sp.killed and sp.killed <= bad                     # only ransomware processes are killed ...
        assert len(sp.killed) >= len(bad) - 1                     # ... and (nearly) all of them
    else:
        assert not sp.killed
