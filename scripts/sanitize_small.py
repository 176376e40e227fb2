"""Small invocations of every kernel for compute-sanitizer (memcheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nerrf_b200 import graph as G
from nerrf_b200.ai.models import GraphSAGE_T, lstm
from nerrf_b200.ai.planner import mcts, rewards
t = lambda a: torch.from_numpy(a).cuda()
for hub in ("src", "dst"):
    g = G.synthetic_graph(N=3000, E=60000, seed=3, hub=hub)
    for algo in ("umma", "ffma"):
        m = GraphSAGE_T(32, 128, 3, algo=algo).cuda()
        h, sc, el = m(t(g.x), t(g.rowptr), t(g.col), t(g.ew), return_edge_logits=True)
        print(hub, algo, float(sc.mean()))
rng = np.random.default_rng(0)
for A in (100, 1500, 4096):
    act = rewards.Actions(rng.beta(0.5, 0.5, A), rng.lognormal(0.7, 1.0, A), np.ones(A))
    r = mcts.search(act, None, 256, 20, 1, iterations=6)
    s = rewards.score(rewards.pack_states(rng.random((9, A)) < 0.3), act)
    print("mcts", A, r.best, float(s.mean()))
mdl = lstm.LSTMScorer().cuda()
print("lstm", mdl(torch.randn(5, 12, 16).cuda(), torch.tensor([12, 1, 5, 9, 3]).cuda()).mean().item())
torch.cuda.synchronize(); print("sanitize_small done")
