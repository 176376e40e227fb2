import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nerrf_b200 import graph as G
from nerrf_b200.ai.models import GraphSAGE_T
from oracle import sage_ref as S
for F in (128, 32):
    for (N, E, hubs) in ((200, 5000, [3]), (300, 9000, [3, 150]), (5000, 60000, None)):
        rng = np.random.default_rng(N)
        if hubs is None:
            g = G.synthetic_graph(N=N, E=E, seed=5, hub="dst", f_in=F)
        else:
            src = rng.integers(0, N, E); dst = rng.choice(hubs, E)
            t = rng.random(E).astype(np.float32) * 60; conf = np.ones(E, np.float32)
            rowptr, col, ew = G.csr_from_edges(src, dst, t, conf, N)
            g = G.TemporalGraph(rowptr, col, ew, rng.standard_normal((N, F)).astype(np.float32), {})
        model = GraphSAGE_T(F, 128, 1, algo="umma").cuda()
        tt = lambda a: torch.from_numpy(a)
        out = model.layer_forward(0, tt(g.x).cuda(), tt(g.rowptr).cuda(), tt(g.col).cuda(), tt(g.ew).cuda(), relu=False).cpu()
        W, b = model.oracle_params()["layers"][0]
        want = S.layer(tt(g.x), tt(g.rowptr), tt(g.col), tt(g.ew), W, b, relu=False)
        err = (out - want).abs().max(1).values
        deg = np.diff(g.rowptr)
        bad = np.nonzero(err.numpy() > 1e-3)[0]
        print(f"F={F} N={N} E={E}: long rows {np.nonzero(deg > 512)[0][:8].tolist()} degs {deg[deg > 512][:8].tolist()} | bad rows {bad[:10].tolist()} (deg {deg[bad][:10].tolist()}) max err {float(err.max()):.3e}")
        for r in bad[:2]:
            print("   row", r, "out", out[r, :3].tolist(), "want", want[r, :3].tolist())
