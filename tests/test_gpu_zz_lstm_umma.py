"""Opt-in tensor-core LSTM path (NERRF_LSTM_ALGO=umma, csrc/lstm_umma.cu) against the oracle, through the same C-ABI
entry point (needs a B200).  Shapes: one partial tile, two tiles with a ragged tail, three tiles -- ragged lengths."""
import pytest
import torch

from nerrf_b200.ai.models.lstm import LSTMScorer
from oracle import lstm_ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,T", [(6, 20), (130, 12), (300, 33)])
def test_umma_path_matches_oracle_and_ffma_path(B, T, monkeypatch):
    torch.manual_seed(B)
    m = LSTMScorer().cuda()
    seq = torch.randn(B, T, 16); ln = torch.randint(1, T + 1, (B,)); ln[0] = T
    want = lstm_ref.forward(m.oracle_params(), seq, ln)
    monkeypatch.delenv("NERRF_LSTM_ALGO", raising=False)
    ffma = m(seq.cuda(), ln.cuda()).cpu()
    monkeypatch.setenv("NERRF_LSTM_ALGO", "umma")
    umma = m(seq.cuda(), ln.cuda()).cpu()
    assert float((ffma - want).abs().max()) < 1e-5
    assert float((umma - want).abs().max()) < 1e-5          # bf16 x 3 split: fp32-equivalent (measured 1.2e-7)
