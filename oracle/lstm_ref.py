"""lstm.forward oracle (plain PyTorch CPU fp32, explicit cell loop; test infra).

Spec source in the reference (prose only; ai/models/lstm.py is named in README.md:73 but does
not exist):
  * "Bidirectional LSTM (256 hidden, 2 layers)", input "last 100 events per file"
                                                docs/content/docs/architecture.mdx:55-59
  * outputs encrypt_probability, ransomware_score   docs/content/docs/threat-model.mdx:191-203
Frozen spec v0: torch.nn.LSTM semantics -- gate order (i, f, g, o),
    i,f,o = sigmoid, g = tanh,  c' = f*c + i*g,  h' = o*tanh(c'),
two biases per layer/direction (b_ih + b_hh), batch_first input [B, T, D_in], valid steps are
t < len[b] (packed-sequence semantics: the forward direction's final state is taken at
t = len-1, the backward direction starts at t = len-1 and ends at t = 0; padded outputs are 0).
Head: [h_fwd_final || h_bwd_final] of the top layer (2H) -> Linear(2H, 2) -> sigmoid
      = (encrypt_probability, ransomware_score).
Pinned against torch.nn.LSTM + pack_padded_sequence in tests/test_oracle_lstm.py.
"""
import torch


def _cell(x_t, h, c, W_ih, W_hh, b_ih, b_hh):
    H = h.shape[1]
    g = x_t @ W_ih.t() + b_ih + h @ W_hh.t() + b_hh
    i = torch.sigmoid(g[:, 0:H]); f = torch.sigmoid(g[:, H:2 * H])
    gg = torch.tanh(g[:, 2 * H:3 * H]); o = torch.sigmoid(g[:, 3 * H:4 * H])
    c2 = f * c + i * gg
    h2 = o * torch.tanh(c2)
    return h2, c2


def run_direction(x, lengths, W_ih, W_hh, b_ih, b_hh, reverse):
    """x [B,T,D] -> (out [B,T,H] zero at padded steps, h_final [B,H])."""
    B, T, _ = x.shape
    H = W_hh.shape[1]
    h = torch.zeros(B, H, dtype=x.dtype); c = torch.zeros(B, H, dtype=x.dtype)
    out = torch.zeros(B, T, H, dtype=x.dtype)
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        valid = (lengths > t)[:, None]
        h2, c2 = _cell(x[:, t], h, c, W_ih, W_hh, b_ih, b_hh)
        h = torch.where(valid, h2, h); c = torch.where(valid, c2, c)
        out[:, t] = torch.where(valid, h2, torch.zeros_like(h2))
    return out, h


def forward(params, seq, lengths):
    """params: {'lstm': [[(W_ih,W_hh,b_ih,b_hh) fwd, (...) bwd] per layer], 'head_W' [2,2H], 'head_b' [2]}
    seq [B,T,D_in] fp32, lengths int [B] -> probs [B,2]."""
    x = seq
    lengths = lengths.to(torch.int64)
    hf = hb = None
    for (fw, bw) in params["lstm"]:
        of, hf = run_direction(x, lengths, *fw, reverse=False)
        ob, hb = run_direction(x, lengths, *bw, reverse=True)
        x = torch.cat([of, ob], dim=2)
    feat = torch.cat([hf, hb], dim=1)
    return torch.sigmoid(feat @ params["head_W"].t() + params["head_b"])


def make_params(in_dim=16, hidden=256, num_layers=2, seed=3):
    """nn.LSTM default init U(-1/sqrt(H), 1/sqrt(H)) from a seeded generator."""
    g = torch.Generator().manual_seed(seed)
    k = 1.0 / hidden ** 0.5

    def u(*shape):
        return (torch.rand(*shape, generator=g) * 2 - 1) * k

    layers = []
    D = in_dim
    for _ in range(num_layers):
        dirs = []
        for _d in range(2):
            dirs.append((u(4 * hidden, D), u(4 * hidden, hidden), u(4 * hidden), u(4 * hidden)))
        layers.append(dirs)
        D = 2 * hidden
    kh = 1.0 / (2 * hidden) ** 0.5
    return {"lstm": layers,
            "head_W": (torch.rand(2, 2 * hidden, generator=g) * 2 - 1) * kh,
            "head_b": (torch.rand(2, generator=g) * 2 - 1) * kh}


def to_nn_lstm(params, in_dim, hidden):
    """Build a torch.nn.LSTM carrying the same weights (used only to pin this oracle)."""
    L = len(params["lstm"])
    m = torch.nn.LSTM(in_dim, hidden, num_layers=L, bidirectional=True, batch_first=True)
    with torch.no_grad():
        for l, (fw, bw) in enumerate(params["lstm"]):
            for sfx, ws in (("", fw), ("_reverse", bw)):
                getattr(m, f"weight_ih_l{l}{sfx}").copy_(ws[0])
                getattr(m, f"weight_hh_l{l}{sfx}").copy_(ws[1])
                getattr(m, f"bias_ih_l{l}{sfx}").copy_(ws[2])
                getattr(m, f"bias_hh_l{l}{sfx}").copy_(ws[3])
    return m
