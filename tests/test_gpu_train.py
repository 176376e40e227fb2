"""Weights trained by ai.train (autograd, CPU) drive the CUDA inference path: same scores, same AUC, and the
pipeline plans from MODEL confidence alone (needs a B200)."""
import numpy as np
import pytest
import torch

from nerrf_b200 import pipeline
from nerrf_b200.ai import train as T
from nerrf_b200.ai.models import GraphSAGE_T
from nerrf_b200.ai.models.lstm import LSTMScorer

pytestmark = pytest.mark.gpu


def test_trained_weights_through_the_cuda_kernels():
    torch.manual_seed(0)
    model, scorer = GraphSAGE_T(32, 128, 2), LSTMScorer()
    T.train(model, scorer, T.toy_set(range(100, 104)), epochs=25, lr=3e-3)
    ex = T.make_example(901, n_files=30, benign_files=40)
    with torch.no_grad():
        want_node = torch.sigmoid(T.sage_node_logits(model, ex["x"], ex["rowptr"], ex["col"], ex["ew"]))
        want_seq = torch.sigmoid(T.lstm_logits(scorer, ex["seq"], ex["lengths"]))
    model.cuda(); scorer.cuda()
    c = lambda k: ex[k].cuda()
    _, score = model(c("x"), c("rowptr"), c("col"), c("ew"))
    probs = scorer(c("seq"), c("lengths"))
    assert float((score.cpu() - want_node).abs().max()) <= 1e-4          # fp32 spec tolerance (north star: 1e-4 rel)
    assert float((probs.cpu() - want_seq).abs().max()) <= 1e-4
    files = ex["is_file"].numpy()
    assert T.roc_auc(score.cpu().numpy()[files], ex["label"].numpy()[files]) >= 0.90
    assert T.roc_auc(probs[:, 0].cpu().numpy(), ex["seq_label"].numpy()) >= 0.90
    # plan from model confidence alone (no labels): exactly the encrypted files are reverted
    g = ex["graph"]
    res = pipeline.run(g, ex["seq"].numpy(), ex["lengths"].numpy(), ex["seq_nodes"], model, scorer, top_a=4096,
                       n_rollouts=1024, depth=40, iterations=8)
    assert sorted(res.plan_nodes) == sorted(np.nonzero(g.meta["label"])[0].tolist())
