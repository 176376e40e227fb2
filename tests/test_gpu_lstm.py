"""lstm.forward CUDA path vs the oracle (needs a B200)."""
import pytest
import torch

from nerrf_b200.ai.models import lstm
from oracle import lstm_ref as LR
from gpu_util import assert_close_fp32

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,T", [(1, 1), (5, 7), (37, 100), (700, 20)])
def test_lstm_parity(B, T):
    torch.manual_seed(B * 1000 + T)
    model = lstm.LSTMScorer(16, 256, 2, seed=3).cuda()
    seq = torch.randn(B, T, 16)
    lengths = torch.randint(1, T + 1, (B,))
    lengths[0] = T
    got = model(seq.cuda(), lengths.cuda())
    want = LR.forward(model.oracle_params(), seq, lengths)
    assert got.shape == (B, 2)
    assert_close_fp32(got, want, rtol=1e-4, atol_rms=1e-5, what=f"lstm B={B} T={T}")


def test_padding_is_ignored_and_module_level_forward():
    model = lstm.LSTMScorer().cuda()
    seq = torch.randn(9, 100, 16, device="cuda")
    lengths = torch.tensor([100, 3, 50, 1, 99, 100, 17, 2, 64], device="cuda")
    a = model(seq, lengths)
    seq2 = seq.clone()
    for b in range(9):
        seq2[b, int(lengths[b]):] = 1e6          # garbage in the padded tail
    assert torch.equal(a, model(seq2, lengths))
    assert torch.equal(a, lstm.forward(seq, lengths, model))
    assert (a > 0).all() and (a < 1).all()
