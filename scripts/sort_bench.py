"""Graph-constructor sort stage: own radix sort vs cub (NERRF_GRAPH_SORT=cub), cfg-2 edge list (10 M edges, 1 M nodes) and a
time-presorted edge list.  usage: [NERRF_GRAPH_SORT=cub] python scripts/sort_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nerrf_b200 import graph as G

dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(1)
for (N, E) in ((1_000_000, 10_000_000), (1_000_000, 2_200_000)):
    src = torch.randint(0, N, (E,), device=dev, generator=gen, dtype=torch.int32)
    dst = torch.randint(0, N, (E,), device=dev, generator=gen, dtype=torch.int32)
    t = torch.rand(E, device=dev, generator=gen) * 60
    conf = torch.ones(E, device=dev)
    for label, t in (("random t", t), ("time-sorted input", torch.sort(t).values)):
        for _ in range(2):
            G.build_csr_device(src, dst, t, conf, N)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        for i in range(5):
            ev[i].record(); G.build_csr_device(src, dst, t, conf, N)
        ev[5].record(); torch.cuda.synchronize()
        ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(5))[2]
        print(f"sort={os.environ.get('NERRF_GRAPH_SORT', 'own')} N={N} E={E} {label}: build_csr {ms:.3f} ms = {E / ms / 1e6:.2f} G edges/s", flush=True)
