"""cfg-2 size graph with hub DESTINATIONS (dst = floor(N u^3): rows with up to ~1e5 in-edges): forward time."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nerrf_b200.ai.models import GraphSAGE_T
from nerrf_b200.graph import synthetic_graph
g = synthetic_graph(hub="dst")
print("max in-degree", int(np.diff(g.rowptr).max()), "rows > 512:", int((np.diff(g.rowptr) > 512).sum()))
t = lambda a: torch.from_numpy(a).cuda()
x, rp, col, ew = t(g.x), t(g.rowptr), t(g.col), t(g.ew)
model = GraphSAGE_T(32, 128, 3).cuda()
for _ in range(2): h, sc = model(x, rp, col, ew)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): h, sc = model(x, rp, col, ew)
e1.record(); torch.cuda.synchronize()
print("hub=dst forward ms:", e0.elapsed_time(e1) / 5, "finite:", bool(torch.isfinite(h).all()))
