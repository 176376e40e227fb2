"""ai.train: the autograd restatement equals the oracle spec, roc_auc equals scikit-learn's,
and a short CPU training run passes the reference's M2 gate (ROC-AUC >= 0.90, ROADMAP.md:26,62-69) on held-out
simulated traces AND on the graphs of the reference's own m0/m1 traces (tests/golden/golden_m1_graph.npz)."""
import os

import numpy as np
import pytest
import torch

from nerrf_b200.ai import train as T
from nerrf_b200.ai.models import GraphSAGE_T
from nerrf_b200.ai.models.lstm import LSTMScorer
from oracle import lstm_ref, sage_ref


def test_roc_auc_matches_sklearn_including_ties():
    from sklearn.metrics import roc_auc_score
    rng = np.random.default_rng(0)
    for n in (5, 50, 500):
        y = rng.integers(0, 2, n); y[0], y[1] = 0, 1
        s = np.round(rng.normal(size=n) + y, 1)                 # rounding makes ties
        assert abs(T.roc_auc(s, y) - roc_auc_score(y, s)) < 1e-12
    assert np.isnan(T.roc_auc([0.1, 0.2], [1, 1]))


def test_autograd_restatements_equal_the_oracle_spec():
    ex = T.make_example(3, n_files=8, benign_files=6)
    m = GraphSAGE_T(32, 128, 3)
    logit = T.sage_node_logits(m, ex["x"], ex["rowptr"], ex["col"], ex["ew"])
    _, score = sage_ref.forward(m.oracle_params(), ex["x"], ex["rowptr"], ex["col"], ex["ew"])
    assert float((torch.sigmoid(logit).detach() - score).abs().max()) < 1e-6
    s = LSTMScorer()
    p = torch.sigmoid(T.lstm_logits(s, ex["seq"], ex["lengths"]))
    want = lstm_ref.forward(s.oracle_params(), ex["seq"], ex["lengths"])
    assert float((p.detach() - want).abs().max()) < 1e-6
    # gradients reach every parameter of both models
    (logit.sum() + T.lstm_logits(s, ex["seq"], ex["lengths"]).sum()).backward()
    named = [(n, q) for n, q in m.named_parameters() if not n.startswith("edge_")] + list(s.named_parameters())
    assert all(q.grad is not None and torch.isfinite(q.grad).all() for _, q in named), [n for n, q in named if q.grad is None]


@pytest.fixture(scope="module")
def trained():
    torch.manual_seed(0)
    model, scorer = GraphSAGE_T(32, 128, 2), LSTMScorer()
    T.train(model, scorer, T.toy_set(range(100, 104)), epochs=25, lr=3e-3)
    return model, scorer


def test_m2_gate_on_held_out_traces(trained):
    model, scorer = trained
    auc = T.evaluate(model, scorer, T.toy_set(range(900, 903)))
    assert auc["gnn_auc"] >= 0.90 and auc["lstm_auc"] >= 0.90, auc


@pytest.mark.parametrize("name", ["m0", "m1"])
def test_m2_gate_on_the_reference_traces(trained, name):
    model, _ = trained
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_m1_graph.npz"))
    t = torch.from_numpy
    with torch.no_grad():
        logit = T.sage_node_logits(model, t(z[name + "_x_obs"]), t(z[name + "_rowptr"]), t(z[name + "_col"]), t(z[name + "_ew"]))
    files = z[name + "_kind"] == 0
    auc = T.roc_auc(logit.numpy()[files], z[name + "_label"][files])
    assert auc >= 0.90, auc


def test_training_features_carry_no_annotation_slots():
    """the label-defining event kinds are not features (graph.OBSERVABLE_SLOT)."""
    ex = T.make_example(5, n_files=6, benign_files=5)
    assert not ex["x"][:, 5:9].any() and not ex["seq"][:, :, 0:4].any()
    assert ex["label"].sum() == 6


def test_checkpoint_round_trip_into_the_undo_cli_loader(tmp_path):
    from nerrf_b200 import undo
    out = str(tmp_path / "w.pt")
    T.main(["--traces", "2", "--epochs", "2", "--out", out])
    model, scorer = undo.load_models(out, 0)
    ck = torch.load(out, map_location="cpu")
    assert model.num_layers == ck["layers"] == 2
    assert all(torch.equal(v, ck["sage"][k]) for k, v in model.state_dict().items())
    assert all(torch.equal(v, ck["lstm"][k]) for k, v in scorer.state_dict().items())
    with pytest.raises(SystemExit, match="no weights"):
        undo.load_models(None, 0)
