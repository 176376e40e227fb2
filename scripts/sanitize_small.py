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

# ---- round 2 kernels: backward, device interning + radix sort + features + sequences (one stream window), plan commit,
# pipelined host session
from nerrf_b200 import stream, ingest
from nerrf_b200.ai import autograd as AG
from nerrf_b200.ai.models.graphsage_t import HostSession
g = G.synthetic_graph(N=1500, E=20000, seed=4)
m = GraphSAGE_T(32, 128, 2).cuda()
xg = t(g.x).requires_grad_()
tg = AG.TrainGraph(t(g.rowptr), t(g.col), t(g.ew))
AG.sage_node_logits(m, xg, tg).sum().backward()
print("bwd", float(xg.grad.abs().mean()), float(m.weights[0].grad.abs().mean()))
cols, _ = stream.fleet_columns(40, 3, seed=2)
ds = stream.DeviceStream(cols, "cuda")
t0, t1 = ds.span()
gw = ds.window_graph(t0 + 5, t0 + 65, 60.0)
files = np.nonzero(gw.meta["node_kind"] == 0)[0][:300]
sq, ln = ds.sequences_device(gw, files)
print("stream window", gw.num_nodes, gw.num_edges, float(sq.sum()), int(ln.sum()))
for merge in (True, False):
    r = ingest.intern_nodes_device(cols, np.argsort(cols.timestamp, kind="stable"), merge)
    print("intern", merge, r[3].shape[0])
rng = np.random.default_rng(1)
A = 600
guard = np.full(A, -1, np.int32); guard[16:] = rng.integers(0, 16, A - 16)
act = rewards.Actions(rng.beta(0.5, 0.5, A), rng.lognormal(0.7, 1.0, A), np.ones(A), guard=guard)
pl = mcts.plan(act, n_rollouts=256, depth=16, iterations=4, commit_per_search=32, max_steps=100)
print("plan", len(pl.actions))
sess = HostSession(m, 2000, 30000)
pin = lambda a: torch.from_numpy(a).pin_memory()
sc_out = [torch.empty(g.num_nodes).pin_memory() for _ in range(2)]
tk = [sess.submit(pin(g.x), pin(g.rowptr), pin(g.col), pin(g.ew), sc_out[i]) for i in range(2)]
[sess.wait(k) for k in tk]; sess.close()
print("session", float(sc_out[0].mean()))
torch.cuda.synchronize(); print("sanitize_small round-2 part done")
