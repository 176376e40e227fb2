"""One window of the device-resident fleet stream (target for ncu: intern_*, rs_*, node_* kernels)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nerrf_b200 import stream
cols, _ = stream.fleet_columns(10400, 40, seed=5)
ds = stream.DeviceStream(cols, "cuda")
t0, t1 = ds.span()
for _ in range(2):
    g = ds.window_graph(t0 + 30.0, t0 + 90.0, 60.0)
torch.cuda.synchronize()
a = time.perf_counter(); g = ds.window_graph(t0 + 30.0, t0 + 90.0, 60.0); torch.cuda.synchronize()
print("window graph: %d events -> %d nodes, %d edges in %.2f ms" % (g.meta["events"], g.num_nodes, g.num_edges, (time.perf_counter() - a) * 1e3))
import numpy as np
files = np.nonzero(g.meta["node_kind"] == 0)[0][:4096]
sq, ln = ds.sequences_device(g, files); torch.cuda.synchronize()
from nerrf_b200.ai.planner import mcts, rewards
rng = np.random.default_rng(1)
A = 4096
guard = np.full(A, -1, np.int32); guard[32:] = rng.integers(0, 32, A - 32)
act = rewards.Actions(rng.beta(0.5, 0.5, A), rng.lognormal(0.7, 1.0, A), np.ones(A), guard=guard)
pl = mcts.plan(act, n_rollouts=1024, depth=32, iterations=8, commit_per_search=64, max_steps=128)
print("sequences", tuple(sq.shape), "plan", len(pl.actions))
