"""Run the cfg-2 GraphSAGE-T layers a few times (target for ncu).  usage: profile_sage.py [algo] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nerrf_b200.ai.models import GraphSAGE_T
from nerrf_b200.graph import synthetic_graph

algo = sys.argv[1] if len(sys.argv) > 1 else "auto"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = synthetic_graph()
t = lambda a: torch.from_numpy(a).cuda()
x, rp, col, ew = t(g.x), t(g.rowptr), t(g.col), t(g.ew)
model = GraphSAGE_T(32, 128, 3, algo=algo).cuda()
for _ in range(reps):
    h, sc = model(x, rp, col, ew)
torch.cuda.synchronize()
print("ok", float(sc.mean()))
