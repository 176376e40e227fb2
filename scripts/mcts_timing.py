"""MCTS timing (cfg 3): device time per search at T=16 and T=64, host-session e2e."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nerrf_b200.ai.planner import mcts
from nerrf_b200.ai.planner.rewards import Actions
rng = np.random.default_rng(2)
A, R, D = 1024, 4096, 50
act = Actions(rng.beta(0.5, 0.5, A), rng.lognormal(np.log(2.0), 1.0, A), rng.choice([1.0, 10.0, 100.0], A, p=[.9, .09, .01]))
for T in (16, 64):
    ctx = mcts.SearchContext(act, R, D, T)
    ctx.search(0); torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(7)]
    for i, (a, b) in enumerate(evs):
        a.record(); ctx.launch(i); b.record()
    torch.cuda.synchronize()
    ms = float(np.median([a.elapsed_time(b) for a, b in evs]))
    sess = mcts.HostSession(A, T, R)
    mcts.search(act, None, R, D, 0, iterations=T, host_call=sess)
    t0 = time.perf_counter()
    for i in range(5): mcts.search(act, None, R, D, i, iterations=T, host_call=sess)
    e2e = (time.perf_counter() - t0) / 5
    print(f"T={T}: {ms:.3f} ms/search = {ms / T * 1e3:.1f} us/iteration, {R * T / ms / 1e3:.1f} M rollouts/s; host session {e2e * 1e3:.3f} ms = {R * T / e2e / 1e6:.1f} M rollouts/s", flush=True)
ctx = mcts.SearchContext(act, R, D, 64)
ctx.search(0); torch.cuda.synchronize()
prof = ctx.ws[:64].view(torch.int32)[8:12].cpu().numpy().astype(np.int64) * 16
print("CTA0 cycles per iteration: select %.0f rollouts %.0f barrier %.0f backup %.0f (sum %.0f = %.1f us at 1.9 GHz)" % (*(prof / 64), prof.sum() / 64, prof.sum() / 64 / 1900))
