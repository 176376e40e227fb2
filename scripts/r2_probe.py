"""Box probe (round 2): host CPU, C sage oracle on the full cfg-2 graph with all cores (the honest CPU arm)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import c_sage, sage_ref as S
from nerrf_b200.graph import synthetic_graph
model = [l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")]
print("cpu:", model[0], "x", len(model), "os.cpu_count", os.cpu_count())
t0 = time.perf_counter(); g = synthetic_graph(); print("graph gen s", time.perf_counter() - t0)
P = S.make_params(32, 128, 3, seed=1)
out = {}
for th in (16, 32, 64, 128, 256):
    if th > (os.cpu_count() or 1): continue
    os.environ["OMP_NUM_THREADS"] = str(th)
    # threads are fixed at first omp use per process: re-exec per setting
    import subprocess
    r = subprocess.run([sys.executable, "-c", f"""
import sys, time, os
sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
from oracle import c_sage, sage_ref as S
from nerrf_b200.graph import synthetic_graph
g = synthetic_graph(); P = S.make_params(32,128,3,seed=1)
f = c_sage.Forward(P, g.x, g.rowptr, g.col, g.ew); f.run()
ts=[]
for _ in range(3):
    t0=time.perf_counter(); f.run(); ts.append(time.perf_counter()-t0)
print(c_sage.threads(), min(ts))
"""], capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS=str(th), OMP_PROC_BIND="spread"))
    print("threads", th, r.stdout.strip(), r.stderr[-300:])
