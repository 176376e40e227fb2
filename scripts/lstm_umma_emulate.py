"""CPU emulation of the DATA FLOW of csrc/lstm_umma.cu (packing, gate-slice permutation, slab layouts, K halves of the
input projection, group/slice recurrence, masking) against oracle/lstm_ref.py.  It mirrors the index arithmetic of the
kernels one to one (same formulas), with exact fp32 math instead of the bf16x3 split, so an indexing mistake in the design
shows up here without a GPU.  Not a test of the PTX."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import lstm_ref  # noqa: E402

LH, SL, UM, NSEQ = 256, 8, 128, 128


def pack_slices(Wt, k_lo, k_real, K_dst):
    """pack_slices_kernel: dst[s][k][j] = Wt[(k_lo+k)*1024 + (j>>5)*256 + s*32 + (j&31)] for k < k_real else 0"""
    dst = np.zeros((SL, K_dst, UM), np.float32)
    for s in range(SL):
        for j in range(UM):
            g = (j >> 5) * LH + s * 32 + (j & 31)
            dst[s, :k_real, j] = Wt[k_lo:k_lo + k_real, g]
    return dst


def pack_bias(bias):
    dst = np.zeros((SL, UM), np.float32)
    for s in range(SL):
        for j in range(UM):
            dst[s, j] = bias[(j >> 5) * LH + s * 32 + (j & 31)]
    return dst


def sage_dense(x, self_rows, nbr_rows, W, b):
    """sage_layer_umma with relu=0 on the identity-shift graph: out[v] = [x[self] || x[nbr] or 0] @ W + b"""
    F = x.shape[1]
    m = x[nbr_rows] if nbr_rows is not None else np.zeros((len(self_rows), F), np.float32)
    return np.concatenate([x[self_rows], m], 1) @ W + b


def emulate(params, seq, lengths):
    B, T, D = seq.shape
    R = B * T
    L = len(params["lstm"])
    slabs = np.zeros((4 * R, UM), np.float32)
    xpad = np.zeros((R, 32), np.float32); xpad[:, :D] = seq.transpose(1, 0, 2).reshape(R, D)   # TIME-MAJOR rows: t*B + b
    hfin = np.zeros((B, 2 * LH), np.float32)
    v = np.arange(R)
    for l in range(L):
        gxa = np.zeros((2 * SL, R, UM), np.float32); gxb = np.zeros_like(gxa)
        whh = np.zeros((2, SL, LH, UM), np.float32)
        for d in range(2):
            W_ih, W_hh, b_ih, b_hh = [p.numpy() for p in params["lstm"][l][d]]
            Wih_t, Whh_t, bias = W_ih.T.copy(), W_hh.T.copy(), b_ih + b_hh           # what nerrf_lstm_forward receives
            bp = pack_bias(bias)
            for kh in range(1 if l == 0 else 2):
                wih = pack_slices(Wih_t, 0, D, 64) if l == 0 else pack_slices(Wih_t, kh * 256, 256, 256)
                for s in range(SL):
                    if l == 0:
                        out = sage_dense(xpad, v, None, wih[s], bp[s])
                    else:
                        x = slabs[kh * 2 * R:]                                         # pointer offset kh*2*R*UM
                        out = sage_dense(x, v, R + v, wih[s], bp[s] if kh == 0 else 0.0)
                    (gxa if kh == 0 else gxb)[d * SL + s] = out
            whh[d] = pack_slices(Whh_t, 0, LH, LH)
        # recurrence: one "group" per (dir, tile); slices are the 8 CTAs
        write_out = l < L - 1
        for d in range(2):
            for b0 in range(0, B, NSEQ):
                nb = min(NSEQ, B - b0)
                h = np.zeros((NSEQ, LH), np.float32); c = np.zeros((NSEQ, LH), np.float32)
                ln = np.zeros(NSEQ, np.int64); ln[:nb] = lengths[b0:b0 + nb]
                for step in range(T):
                    t = T - 1 - step if d else step
                    hn = h.copy()
                    for s in range(SL):
                        rows = t * B + b0 + np.arange(nb)
                        g = np.zeros((NSEQ, UM), np.float32)
                        g[:nb] = gxa[d * SL + s][rows] + (gxb[d * SL + s][rows] if l > 0 else 0.0)
                        g = g + h @ whh[d, s]                                           # gates^T = W_slice . h^T
                        i_, f_, g_, o_ = (g[:, 32 * w:32 * w + 32] for w in range(4))
                        sig = lambda z: 1.0 / (1.0 + np.exp(-z))
                        u = slice(s * 32, s * 32 + 32)
                        c2 = sig(f_) * c[:, u] + sig(i_) * np.tanh(g_)
                        h2 = sig(o_) * np.tanh(c2)
                        valid = (t < ln)[:, None]
                        c[:, u] = np.where(valid, c2, c[:, u]); hn[:, u] = np.where(valid, h2, h[:, u])
                        if write_out:
                            for uu in range(32):
                                feat = d * LH + s * 32 + uu
                                slabs[(feat >> 7) * R + rows, feat & 127] = np.where(valid[:nb, 0], h2[:nb, uu], 0.0)
                    h = hn
                hfin[b0:b0 + nb, d * LH:(d + 1) * LH] = h[:nb]
    z = hfin @ params["head_W"].numpy().T + params["head_b"].numpy()
    return 1.0 / (1.0 + np.exp(-z))


if __name__ == "__main__":
    torch.manual_seed(0)
    params = lstm_ref.make_params(16, 256, 2, seed=5)
    B, T = 5, 7
    seq = torch.randn(B, T, 16); lengths = torch.tensor([7, 1, 4, 7, 2])
    want = lstm_ref.forward(params, seq, lengths).numpy()
    got = emulate(params, seq.numpy(), lengths.numpy())
    print("max |emulation - oracle| =", float(np.abs(got - want).max()))
    assert np.abs(got - want).max() < 1e-5
    print("data flow of lstm_umma.cu matches the oracle")
