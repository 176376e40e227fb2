"""Kernel-time split of one tensor-core LSTM forward (torch.profiler / CUPTI sees the library's kernels too)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from nerrf_b200.ai.models.lstm import LSTMScorer  # noqa: E402

m = LSTMScorer().cuda()
B, T = 4096, 100
seq = torch.randn(B, T, 16, device="cuda"); ln = torch.randint(T // 2, T + 1, (B,), device="cuda")
m(seq, ln); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    m(seq, ln); torch.cuda.synchronize()
rows = {}
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        k = e.name[:70]
        t, n = rows.get(k, (0.0, 0))
        rows[k] = (t + e.device_time_total if hasattr(e, "device_time_total") else t + e.cuda_time_total, n + 1)
for k, (t, n) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:12]:
    print(f"{t / 1e3:9.3f} ms  x{n:<4d} {k}")
