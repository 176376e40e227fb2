"""Pin the LSTM oracle against torch.nn.LSTM + pack_padded_sequence (the canonical third-party
implementation the reference intended: README.md:100 'PyTorch')."""
import torch
from torch.nn.utils.rnn import pack_padded_sequence

from oracle import lstm_ref as LR


def test_matches_nn_lstm_packed():
    torch.manual_seed(0)
    B, T, D, H = 7, 12, 16, 32
    P = LR.make_params(D, H, 2, seed=3)
    seq = torch.randn(B, T, D)
    lengths = torch.tensor([12, 1, 5, 12, 7, 3, 9])
    got = LR.forward(P, seq, lengths)
    m = LR.to_nn_lstm(P, D, H)
    with torch.no_grad():
        packed = pack_padded_sequence(seq, lengths, batch_first=True, enforce_sorted=False)
        _, (hn, _) = m(packed)
        feat = torch.cat([hn[-2], hn[-1]], dim=1)
        want = torch.sigmoid(feat @ P["head_W"].t() + P["head_b"])
    assert torch.allclose(got, want, atol=1e-6)


def test_single_step_cell_by_hand():
    H, D = 2, 1
    W_ih = torch.tensor([[1.0], [0.0], [0.5], [0.0], [2.0], [0.0], [-1.0], [0.0]])    # [4H, D]
    W_hh = torch.zeros(4 * H, H); b = torch.zeros(4 * H)
    x = torch.tensor([[[1.0]]])
    out, h = LR.run_direction(x, torch.tensor([1]), W_ih, W_hh, b, b, reverse=False)
    i0 = torch.sigmoid(torch.tensor(1.0)); g0 = torch.tanh(torch.tensor(2.0)); o0 = torch.sigmoid(torch.tensor(-1.0))
    c0 = i0 * g0
    assert torch.allclose(h[0, 0], o0 * torch.tanh(c0), atol=1e-7)
    # unit 1: i = sig(0) = .5, g = tanh(0) = 0 -> c = 0 -> h = 0
    assert abs(float(h[0, 1])) < 1e-7
