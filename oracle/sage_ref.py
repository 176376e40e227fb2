"""GraphSAGE_T.forward oracle (plain PyTorch CPU fp32; test infra -- see oracle/__init__.py).

Spec source in the reference (prose only; ai/models/GraphSAGE-T.py is named in README.md:73 and
ROADMAP.md:127 but does not exist):
  * "classify edges as normal/attack"                  docs/content/docs/architecture.mdx:49-53
  * node ``anomaly_score`` in [0,1]                     docs/content/docs/architecture.mdx:157
  * 2-way edge probabilities                            docs/content/docs/threat-model.mdx:186-188
  * edge weight = causality confidence, 30-60 s window  docs/content/docs/architecture.mdx:40-42
Where silent: canonical GraphSAGE mean aggregator (Hamilton et al. 2017).

Frozen spec v0:
    m_v  = sum_{e: dst(e)=v} w_e * h_src(e)  /  max(sum_e w_e, 1e-12)      (isolated node -> 0)
    h'_v = ReLU( [h_v || m_v] @ W_l + b_l ),   W_l in R^{2F_l x H}
    node_score_v = sigmoid( h_v . w_n + b_n )
    edge_logit_e = [h_src(e) || h_dst(e)] @ W_e + b_e          (2 classes: attack, normal)
Graph: CSR by destination (rowptr [N+1], col [E] = source ids, ew [E] fp32 temporal weights
``w_e = conf_e * exp(-(t_ref - t_e)/tau)`` precomputed on the host).
This is the "reference ai/ CPU path" that bench.py times as cpu_baseline: index_select gather
+ index_add_ scatter + addmm, all host threads.
"""
import torch


def edge_dst(rowptr):
    rowptr = rowptr.to(torch.int64)
    deg = rowptr[1:] - rowptr[:-1]
    return torch.repeat_interleave(torch.arange(deg.numel(), dtype=torch.int64), deg)


def aggregate(h, rowptr, col, ew, dst=None, chunk=1 << 21, row_begin=0, row_end=None, dtype=None):
    """Weighted segmented mean over CSR rows [row_begin,row_end).  Returns [rows, F]."""
    dtype = dtype or h.dtype
    N = rowptr.numel() - 1
    row_end = N if row_end is None else row_end
    rowptr = rowptr.to(torch.int64)
    e0 = int(rowptr[row_begin]); e1 = int(rowptr[row_end])
    if dst is None:
        dst = edge_dst(rowptr)
    rows = row_end - row_begin
    acc = torch.zeros(rows, h.shape[1], dtype=dtype)
    wsum = torch.zeros(rows, dtype=dtype)
    col = col.to(torch.int64)
    hh = h.to(dtype)
    for s in range(e0, e1, chunk):
        e = min(s + chunk, e1)
        d = dst[s:e] - row_begin
        w = ew[s:e].to(dtype)
        g = hh.index_select(0, col[s:e]) * w[:, None]
        acc.index_add_(0, d, g)
        wsum.index_add_(0, d, w)
    return acc / wsum.clamp_min(1e-12)[:, None]


def layer(h, rowptr, col, ew, W, b, relu=True, dst=None, row_begin=0, row_end=None, dtype=None):
    dtype = dtype or h.dtype
    N = rowptr.numel() - 1
    row_end = N if row_end is None else row_end
    m = aggregate(h, rowptr, col, ew, dst=dst, row_begin=row_begin, row_end=row_end, dtype=dtype)
    z = torch.addmm(b.to(dtype), torch.cat([h[row_begin:row_end].to(dtype), m], dim=1), W.to(dtype))
    return torch.relu(z) if relu else z


def forward(params, x, rowptr, col, ew, edge_logits=False, dtype=None):
    """params: dict with 'layers' = [(W [2F,H], b [H]), ...], 'node_w' [H], 'node_b' [1],
    optional 'edge_W' [2H,2], 'edge_b' [2].  Returns (h, node_score[, edge_logit])."""
    dtype = dtype or x.dtype
    dst = edge_dst(rowptr)
    h = x.to(dtype)
    for (W, b) in params["layers"]:
        h = layer(h, rowptr, col, ew, W, b, relu=True, dst=dst, dtype=dtype)
    score = torch.sigmoid(h @ params["node_w"].to(dtype) + params["node_b"].to(dtype))
    if not edge_logits:
        return h, score
    src = col.to(torch.int64)
    el = torch.cat([h.index_select(0, src), h.index_select(0, dst)], dim=1) @ params["edge_W"].to(dtype) \
        + params["edge_b"].to(dtype)
    return h, score, el


def make_params(in_dim=32, hidden=128, num_layers=3, seed=1, edge_head=True):
    """Xavier-uniform weights from torch.Generator().manual_seed(seed) (SURVEY.md 8d cfg 1)."""
    g = torch.Generator().manual_seed(seed)

    def xavier(fan_in, fan_out):
        a = (6.0 / (fan_in + fan_out)) ** 0.5
        return (torch.rand(fan_in, fan_out, generator=g) * 2 - 1) * a

    layers = []
    F = in_dim
    for _ in range(num_layers):
        W = xavier(2 * F, hidden)
        b = (torch.rand(hidden, generator=g) * 2 - 1) * 0.1
        layers.append((W, b))
        F = hidden
    p = {"layers": layers, "node_w": xavier(hidden, 1)[:, 0].contiguous(),
         "node_b": (torch.rand(1, generator=g) * 2 - 1) * 0.1}
    if edge_head:
        p["edge_W"] = xavier(2 * hidden, 2)
        p["edge_b"] = (torch.rand(2, generator=g) * 2 - 1) * 0.1
    return p
