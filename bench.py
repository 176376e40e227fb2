#!/usr/bin/env python
"""bench.py -- headline benchmark of the NERRF AI hot path on B200 (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): GraphSAGE-T edges/sec (+ MCTS rollouts/sec, reported under "mcts") on the
1M-node / 10M-edge synthetic temporal graph, 3-layer inference (configs[1]).  One "step" = one
full GraphSAGE_T.forward (3 fused gather+aggregate+GEMM layers + node head) over the graph.

  value      graph resident in HBM, CUDA-event timed, whole job (all ranks), max over ranks
  e2e        the same forward through the HOST-buffer C-ABI call (nerrf_sage_session_forward_host):
             pinned host graph -> H2D -> forward -> D2H node scores, every step
  roofline   dominant kernel = the F=128 fused layer (layers 2 and 3): algorithmic bytes
             E*(8+4F) + N*(4+4F+4H) per launch / its CUDA-event duration inside the timed region,
             against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the oracle (oracle/sage_ref.py, the "reference ai/ CPU path"; kind "port" -- the
             reference ships no implementation) on the box's host cores, bounded sample
N > 1: weak scaling -- the graph grows to N x (1M nodes, 10M edges), 1-D edge-block sharded with
row-aligned cuts, one embedding exchange per layer over NCCL (nerrf_b200/dist.py).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_NODES, N_EDGES, F_IN, HIDDEN, LAYERS = 1_000_000, 10_000_000, 32, 128, 3
MCTS_CFG = dict(A=1024, R=4096, D=50, T=64)


def algorithmic_bytes_layer(E, N, F, H=HIDDEN, s_rp=4):
    """col + w + gathered source row per edge; rowptr + self row + output row per node."""
    return E * (8 + 4 * F) + N * (s_rp + 4 * F + 4 * H)


def algorithmic_bytes_forward(E, N):
    b, F = 0, F_IN
    for _ in range(LAYERS):
        b += algorithmic_bytes_layer(E, N, F)
        F = HIDDEN
    return b + 4 * N


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons via NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.002)
        except Exception as e:       # NVML missing: report it rather than fail the bench
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ------------------------------------------------------------------------------------------ CPU arm
_BEST_THREADS = None


def best_cpu_threads(g, params):
    """PyTorch's CPU index_add_/index_select path does not scale to every core of a large host (it got
    SLOWER beyond ~32 threads on the 128-core GPU box), so the CPU baseline is given the thread count at
    which it runs fastest on a small probe (layer 2 on 20k rows); `cores` reports the count actually used."""
    global _BEST_THREADS
    if _BEST_THREADS is not None:
        return _BEST_THREADS
    import torch
    from oracle import sage_ref as S
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu} | {min(ncpu, 8)})
    t = lambda a: torch.from_numpy(a)
    rows = min(20_000, g.num_nodes)
    e_s = int(g.rowptr[rows])
    rp, col, ew = t(g.rowptr)[:rows + 1], t(g.col)[:e_s], t(g.ew)[:e_s]
    dst = S.edge_dst(rp)
    h = torch.randn(g.num_nodes, HIDDEN, generator=torch.Generator().manual_seed(0))
    W, b = params["layers"][1]
    best, best_dt = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        S.layer(h, rp, col, ew, W, b, dst=dst, row_begin=0, row_end=rows)          # warm
        t0 = time.perf_counter()
        S.layer(h, rp, col, ew, W, b, dst=dst, row_begin=0, row_end=rows)
        dt = time.perf_counter() - t0
        if dt < best_dt:
            best, best_dt = c, dt
    _BEST_THREADS = best
    return best


def cpu_reference_sample(g, params, rows=200_000, threads=None):
    """Time the oracle on a bounded sample: destination rows [0, rows) of EVERY layer, with
    full-size inputs (layer 1 reads x; layers 2/3 read a full-size [N,128] activation), so the
    gather has the real working set.  Returns (edges_per_s, seconds, sampled_edges, threads)."""
    import torch
    from oracle import sage_ref as S
    threads = threads or best_cpu_threads(g, params)
    torch.set_num_threads(threads)
    t = lambda a: torch.from_numpy(a)
    x, rp, col, ew = t(g.x), t(g.rowptr), t(g.col), t(g.ew)
    rows = min(rows, g.num_nodes)
    e_s = int(g.rowptr[rows])
    dst = S.edge_dst(rp[:rows + 1])
    h = torch.relu(torch.randn(g.num_nodes, HIDDEN, generator=torch.Generator().manual_seed(0)))
    t0 = time.perf_counter()
    inp = x
    for (W, b) in params["layers"]:
        out = S.layer(inp, rp[:rows + 1], col[:e_s], ew[:e_s], W, b, dst=dst, row_begin=0, row_end=rows)
        inp = h
    _ = torch.sigmoid(out @ params["node_w"] + params["node_b"])
    dt = time.perf_counter() - t0
    return e_s / dt, dt, e_s, torch.get_num_threads()


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The reference ships
    none, so this is the oracle port on all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from nerrf_b200.graph import synthetic_graph
    from oracle import sage_ref as S
    g = synthetic_graph(N_NODES, N_EDGES)
    params = S.make_params(F_IN, HIDDEN, LAYERS, seed=1)
    rows = 100_000
    vals = []
    for i in range(args.warmup + args.steps):
        eps, dt, es, th = cpu_reference_sample(g, params, rows=rows)
        if i >= args.warmup:
            vals.append((eps, dt))
    eps = float(np.mean([v[0] for v in vals])); dt = float(np.mean([v[1] for v in vals]))
    sample = f"destination rows [0,{rows}) of all {LAYERS} layers ({es} edges, {100.0 * es / N_EDGES:.1f}% of the graph), full-size inputs"
    line = {"impl": "reference", "metric": "graphsage_t_edges_per_sec", "value": eps, "unit": "edges/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": eps, "unit": "edges/s", "cores": th, "kind": "port", "sample": sample},
            "e2e": {"value": eps, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    return {"workload": f"GraphSAGE-T {LAYERS}-layer inference, {n_gpus}x(1M-node / 10M-edge) synthetic temporal graph "
                        f"(PCG64 seed 20250115, src=floor(N*u^3), dst~U), F_in={F_IN}, H={HIDDEN}",
            "nodes": N_NODES * n_gpus, "edges": N_EDGES * n_gpus, "layers": LAYERS,
            "parallelism": "single GPU" if n_gpus == 1 else f"1-D edge-block shards x{n_gpus}, one embedding exchange per layer",
            "l2": "inputs exceed L2 (graph 0.2 GB + activations 0.5 GB/layer vs 126 MB); no flush"}


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from nerrf_b200.graph import synthetic_graph
    from nerrf_b200.ai.models import GraphSAGE_T
    from nerrf_b200.ai.models.graphsage_t import HostSession

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, W = args.steps, max(args.warmup, 3)

    model = GraphSAGE_T(F_IN, HIDDEN, LAYERS, algo=args.algo).to(dev)
    if world == 1:
        g = synthetic_graph(N_NODES, N_EDGES)
        pin = lambda a: torch.from_numpy(a).pin_memory()
        hx, hrp, hcol, hew = pin(g.x), pin(g.rowptr), pin(g.col), pin(g.ew)
        x, rp, col, ew = (t.to(dev) for t in (hx, hrp, hcol, hew))
        N, E = g.num_nodes, g.num_edges
        h_a = torch.empty(N, HIDDEN, device=dev); h_b = torch.empty(N, HIDDEN, device=dev)

        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(LAYERS + 1)] for _ in range(K)]
        score = torch.empty(N, device=dev)

        def step(i=None):
            inp, bufs = x, (h_a, h_b)
            if i is not None: ev[i][0].record()
            for l in range(LAYERS):
                out = bufs[l & 1]
                # the node head is fused into the last layer's epilogue
                model.layer_forward(l, inp, rp, col, ew, out=out, score_out=score if l == LAYERS - 1 else None,
                                    reuse_long_scan=l > 0)
                if i is not None: ev[i][l + 1].record()
                inp = out
            return score

        for _ in range(W):
            step()
        torch.cuda.synchronize()
        sampler = ClockSampler(physical_gpu_index(local_rank)); sampler.start()
        t_start = torch.cuda.Event(enable_timing=True); t_end = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t_start.record()
        for i in range(K):
            step(i)
        t_end.record()
        torch.cuda.synchronize()
        clocks = sampler.result()
        total_ms = t_start.elapsed_time(t_end)
        layer_ms = np.array([[ev[i][l].elapsed_time(ev[i][l + 1]) for l in range(LAYERS)] for i in range(K)])
        ms_per_step = total_ms / K
        value = E / (ms_per_step * 1e-3)
        gpu_launches = K * LAYERS

        # roofline of the dominant kernel: the F=128 fused layer (layers 2..L)
        peak, peak_src = measured_peaks()
        dom_ms = float(layer_ms[:, 1:LAYERS - 1].mean()) if LAYERS > 2 else float(layer_ms[:, 1:].mean())   # middle F=128 layer(s): no fused head
        dom_bytes = algorithmic_bytes_layer(E, N, HIDDEN)
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("sage_layer_F128_dram_bytes_per_launch")
        roofline = {"bound": "hbm", "kernel": "fused GraphSAGE-T layer F=128 (gather+aggregate+GEMM)",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes, "kernel_ms": dom_ms,
                    "per_layer_ms": [float(v) for v in layer_ms.mean(0)],
                    "forward": {"algorithmic_bytes": algorithmic_bytes_forward(E, N),
                                "achieved": algorithmic_bytes_forward(E, N) / (ms_per_step * 1e-3) / 1e9,
                                "frac": algorithmic_bytes_forward(E, N) / (ms_per_step * 1e-3) / 1e9 / peak}}

        # e2e through the host-buffer C-ABI call
        sess = HostSession(model, N, E)
        score_host = torch.empty(N).pin_memory()
        for _ in range(2):
            sess.forward(hx, hrp, hcol, hew, score_host)
        t0 = time.perf_counter()
        for _ in range(K):
            sess.forward(hx, hrp, hcol, hew, score_host)
        e2e_s = (time.perf_counter() - t0) / K
        h2d = int(hx.numel() * 4 + hrp.numel() * 4 + hcol.numel() * 4 + hew.numel() * 4)
        e2e = {"value": E / e2e_s, "unit": "edges/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(N * 4),
               "ms_per_step": e2e_s * 1e3, "api": "nerrf_sage_session_forward_host (HostSession.forward)"}
        sess.close()
        # the device path and the host path agree
        assert torch.equal(score_host, step().cpu()), "host-session scores differ from the device path"

        mcts_info = run_mcts_bench(dev, args)
        lstm_info = run_lstm_bench(dev)
        graph_info = run_graph_build_bench(dev, rp, col)
        cpu = None
        if not args.no_cpu_baseline:
            from oracle import sage_ref as S
            eps, dt, es, th = cpu_reference_sample(g, S.make_params(F_IN, HIDDEN, LAYERS, seed=1), rows=200_000)
            cpu = {"value": eps, "unit": "edges/s", "cores": th, "kind": "port", "seconds": dt,
                   "sample": f"oracle/sage_ref.py, destination rows [0,200000) of all {LAYERS} layers ({es} edges), full-size inputs"}
        lstm_info["umma"] = run_lstm_umma_bench(dev, lstm_info)          # last GPU work of the run (opt-in path)
        line = {"metric": "graphsage_t_edges_per_sec", "value": value, "unit": "edges/s", "n_gpus": 1, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": workload_config(1), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": gpu_launches, "clocks": clocks, "mcts": mcts_info, "lstm": lstm_info, "graph_build": graph_info, "algo": args.algo}
        print(json.dumps(line), flush=True)
        return

    # ---------------------------------------------------------------- N > 1 (weak scaling, sharded)
    from nerrf_b200 import dist as nd
    line = nd.bench_sharded(model, args, world, rank, local_rank, dev, workload_config, algorithmic_bytes_layer,
                            measured_peaks, ClockSampler, physical_gpu_index, run_mcts_bench)
    if rank == 0:
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def run_mcts_bench(dev, args, seed=0):
    """cfg 3: A=1024, R=4096 leaf-parallel rollouts, depth 50, T iterations; rollouts/s on this GPU."""
    import torch
    from nerrf_b200.ai.planner import mcts
    from nerrf_b200.ai.planner.rewards import Actions
    rng = np.random.default_rng(2)
    A, R, D, T = (MCTS_CFG[k] for k in "ARDT")
    act = Actions(rng.beta(0.5, 0.5, A), rng.lognormal(np.log(2.0), 1.0, A),
                  rng.choice([1.0, 10.0, 100.0], A, p=[.9, .09, .01]))
    ctx = mcts.SearchContext(act, R, D, T, device=dev)
    ctx.search(seed)                                                       # warm-up
    torch.cuda.synchronize()
    # device time of the search itself (memsets + persistent kernel), inputs resident
    reps = 5
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for i in range(reps):
        evs[i][0].record(); lo_inv = ctx.launch(seed + i); evs[i][1].record()
    torch.cuda.synchronize()
    ms = float(np.median([a.elapsed_time(b) for a, b in evs]))
    r = ctx.fetch(*lo_inv)
    # through the public call incl. result read-back, and through the host-buffer C-ABI entry
    t0 = time.perf_counter()
    for i in range(reps):
        ctx.search(seed + i)
    api_s = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    mcts.search(act, None, R, D, seed, iterations=T, host_call=True)
    e2e_s = time.perf_counter() - t0
    cpu = None
    if int(os.environ.get("WORLD_SIZE", "1")) == 1 and (args is None or not getattr(args, "no_cpu_baseline", False)):
        # the C restatement of the oracle (oracle/c/planner_oracle.c, OpenMP over the rollouts of an iteration; it is
        # bit-identical to the numpy oracle, tests/test_oracle_c.py) on all host threads: the same search, 5 times
        from oracle import c_oracle
        c_oracle.search(act.p, act.size, act.cost, R=R, D=D, T=2, seed=seed)            # build + warm
        t0 = time.perf_counter()
        for i in range(5):
            c_oracle.search(act.p, act.size, act.cost, R=R, D=D, T=T, seed=seed + i)
        dt = (time.perf_counter() - t0) / 5
        cpu = {"value": T * R / dt, "unit": "rollouts/s", "cores": os.cpu_count(), "kind": "port", "seconds": dt,
               "sample": f"oracle/c/planner_oracle.c (OpenMP over the {R} rollouts of an iteration), the full {T}-iteration search x5"}
    return {"metric": "mcts_rollouts_per_sec", "value": R * T / (ms * 1e-3), "unit": "rollouts/s", "ms_per_search": ms,
            "cpu_baseline": cpu,
            "api_value": R * T / api_s, "e2e_value": R * T / e2e_s,
            "config": {"actions": A, "rollouts_per_iteration": R, "depth": D, "iterations": T}, "best_action": r.best,
            "note": "value: device time (CUDA events) of one search, inputs resident; api_value: SearchContext.search incl. "
                    "result read-back; e2e_value: nerrf_mcts_search_host (alloc + H2D + search + D2H inside the call)"}


def run_lstm_bench(dev, B=4096, T=100):
    """BiLSTM(256 x 2 layers) over B candidate-file sequences of T=100 events."""
    import torch
    from nerrf_b200.ai.models import lstm
    model = lstm.LSTMScorer().to(dev)
    seq = torch.randn(B, T, 16, device=dev)
    lengths = torch.randint(T // 2, T + 1, (B,), device=dev)
    model(seq, lengths); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); model(seq, lengths); model(seq, lengths); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    flops = B * T * 2.0 * (2 * 1024 * (16 + 256) + 2 * 1024 * (512 + 256))      # both directions, both layers
    info = {"metric": "lstm_sequences_per_sec", "value": B / (ms * 1e-3), "unit": "sequences/s", "ms": ms,
            "tflops_fp32": flops / (ms * 1e-3) / 1e12, "config": {"batch": B, "T": T, "hidden": 256, "layers": 2},
            "algo": "ffma (default)"}
    return info


def run_lstm_umma_bench(dev, default_info, B=4096, T=100):
    """Opt-in tensor-core LSTM path (csrc/lstm_umma.cu, NERRF_LSTM_ALGO=umma), same entry point and shape as run_lstm_bench.
    Called LAST, after every reported number has been measured: whatever happens here cannot cost the bench line."""
    import torch
    from nerrf_b200.ai.models import lstm
    try:
        model = lstm.LSTMScorer().to(dev)
        gen = torch.Generator(device=dev).manual_seed(11)
        seq = torch.randn(B, T, 16, device=dev, generator=gen)
        lengths = torch.randint(T // 2, T + 1, (B,), device=dev, generator=gen)
        ref = model(seq, lengths)
        os.environ["NERRF_LSTM_ALGO"] = "umma"
        got = model(seq, lengths); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); model(seq, lengths); model(seq, lengths); e1.record(); torch.cuda.synchronize()
        ms_u = e0.elapsed_time(e1) / 2
        return {"value": B / (ms_u * 1e-3), "unit": "sequences/s", "ms": ms_u,
                "speedup_vs_default": default_info["ms"] / ms_u, "max_abs_diff_vs_default": float((got - ref).abs().max())}
    except Exception as e:
        return {"error": str(e)[:200]}
    finally:
        os.environ.pop("NERRF_LSTM_ALGO", None)


def run_graph_build_bench(dev, rowptr, col, reps=3):
    """Device graph constructor: the workload's own edges as a shuffled edge list -> CSR."""
    import torch
    from nerrf_b200.graph import build_csr_device, WINDOW
    N, E = rowptr.numel() - 1, col.numel()
    gen = torch.Generator(device=dev).manual_seed(7)
    dst = torch.repeat_interleave(torch.arange(N, device=dev, dtype=torch.int32), (rowptr[1:] - rowptr[:-1]).long())
    perm = torch.randperm(E, generator=gen, device=dev)
    src, dst = col[perm].contiguous(), dst[perm].contiguous()
    t = torch.rand(E, generator=gen, device=dev) * WINDOW
    conf = 0.5 + 0.5 * torch.rand(E, generator=gen, device=dev)
    del perm
    rp2, _, _ = build_csr_device(src, dst, t, conf, N)
    assert torch.equal(rp2, rowptr), "device constructor rowptr differs"
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        build_csr_device(src, dst, t, conf, N)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    alg = 24.0 * E + 4.0 * (N + 1)          # read src,dst,t,conf; write col,ew; write rowptr
    return {"metric": "graph_build_edges_per_sec", "value": E / (ms * 1e-3), "unit": "edges/s", "ms": ms,
            "algorithmic_gbps": alg / (ms * 1e-3) / 1e9, "config": {"nodes": N, "edges": E, "sort": "cub radix, 52-bit key"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--algo", default="auto", choices=["auto", "ffma", "umma", "umma2"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "p2p-all", "multicast", "allgather", "broadcast", "allreduce"],
                    help="N>1: per-layer embedding exchange: p2p = fused into the layer kernel (epilogue stores to "
                         "peer-mapped buffers over NVLink, only the rows a peer references); p2p-all = same, every row "
                         "to every peer; multicast = NVSwitch multimem.st; allgather / broadcast / allreduce = NCCL")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
