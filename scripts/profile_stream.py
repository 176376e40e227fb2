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
