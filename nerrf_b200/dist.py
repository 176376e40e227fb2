"""Multi-GPU paths: one process per GPU, torch.distributed for the plumbing.

GraphSAGE-T   1-D EDGE-BLOCK shards over the CSR-by-destination edge array with ROW-ALIGNED cuts
              (graph.edge_balanced_row_cuts): rank g owns destination rows [cuts[g], cuts[g+1]) and
              therefore the contiguous edge block [rowptr[cuts[g]], rowptr[cuts[g+1]]).  Every rank
              keeps the full node-embedding matrix of the current layer; after each layer the ranks
              exchange the rows they produced (one collective per layer):
                exchange="allgather"  equal row blocks: ONE in-place all-gather per layer (half the traffic of
                                      an all-reduce, no fp32 re-association); falls back to "broadcast" when
                                      the blocks are uneven
                exchange="broadcast"  all-gather-v as one broadcast per owner (edge-balanced uneven cuts)
                exchange="allreduce"  the north star's contract: zero outside the owned rows, SUM
              The last layer's node scores stay sharded (each rank returns its rows).
MCTS          root parallelism, no collective on the data path: rank g runs an independent tree with
              seed + g; the [A]-sized root statistics are summed in rank order afterwards.

The compute is injected (`layer_fn`), so the same sharding / exchange logic runs on CPU under gloo
in tests (with the oracle as the layer) and on GPUs under NCCL (with the CUDA kernels).
"""
from __future__ import annotations

import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from . import graph as G


class Shard:
    """What one rank holds: full rowptr, its edge block of col / ew (indexed from edge_base)."""

    def __init__(self, rowptr, col, ew, rank, world, device=None, cut="edges"):
        """cut="edges": edge-balanced row-aligned cuts (skewed graphs); cut="rows": equal row blocks (when the
        in-degree is uniform these are edge-balanced too, and the exchange can be one in-place all-gather)."""
        rp = rowptr.cpu().numpy() if isinstance(rowptr, torch.Tensor) else np.asarray(rowptr)
        n = rp.shape[0] - 1
        if cut == "rows":
            self.cuts = np.array([(n * g) // world for g in range(world + 1)], dtype=np.int64)
        else:
            self.cuts = G.edge_balanced_row_cuts(rp, world)
        self.rank, self.world = rank, world
        self.row_begin, self.row_end = int(self.cuts[rank]), int(self.cuts[rank + 1])
        self.edge_base, self.edge_end = int(rp[self.row_begin]), int(rp[self.row_end])
        dev = device or (rowptr.device if isinstance(rowptr, torch.Tensor) else "cpu")
        t = lambda a: a if isinstance(a, torch.Tensor) else torch.from_numpy(np.asarray(a))
        self.rowptr = t(rowptr).to(dev)
        self.col = t(col)[self.edge_base:self.edge_end].contiguous().to(dev)
        self.ew = t(ew)[self.edge_base:self.edge_end].contiguous().to(dev)
        self.num_nodes = self.rowptr.numel() - 1


class PeerBuffers:
    """Two [N, hidden] embedding buffers per rank, each mapped into every other rank's address space over
    NVLink (torch symmetric memory; CUDA-IPC fallback).  With these the layer kernel's epilogue writes its rows
    straight into the peers' buffers (exchange="p2p": the fused compute + collective form) and the only
    per-layer collective left is a tiny cross-rank barrier."""

    def __init__(self, N, hidden, device, rank, world, group=None, multicast=True):
        self.rank, self.world = rank, world
        self.hdls = None
        self.mc = [0, 0]
        try:
            import torch.distributed._symmetric_memory as symm
            grp = group or dist.group.WORLD
            self.bufs = [symm.empty(N, hidden, dtype=torch.float32, device=device) for _ in range(2)]
            self.hdls = [symm.rendezvous(b, grp) for b in self.bufs]
            self.peers = [[h.get_buffer(p, (N, hidden), torch.float32) for p in range(world) if p != rank] for h in self.hdls]
            self.kind = "symmetric_memory"
            self.mc = [0, 0]
            if multicast:
                try:
                    mc = [int(h.multicast_ptr) for h in self.hdls]
                    if all(mc):
                        self.mc = mc
                        self.kind = "symmetric_memory + NVSwitch multicast"
                except Exception:
                    pass
        except Exception as e:                                      # pragma: no cover - depends on the torch build
            from torch.multiprocessing.reductions import reduce_tensor
            self.bufs = [torch.empty(N, hidden, dtype=torch.float32, device=device) for _ in range(2)]
            self.peers = []
            for b in self.bufs:
                fn, args = reduce_tensor(b)
                gathered = [None] * world
                dist.all_gather_object(gathered, args, group=group)
                self.peers.append([fn(*gathered[p]) for p in range(world) if p != rank])
            self._tok = torch.zeros(1, device=device)
            self.kind = f"cuda_ipc ({type(e).__name__}: {e})"

    def build_need_mask(self, shard: "Shard", N, group=None):
        """uint8 [N]: bit i set <=> the rank behind peer slot i references that row as a SOURCE in its edge block.
        Graph preprocessing (once per graph): each rank marks the sources of its edges, the bitmaps are
        all-gathered, and every rank keeps the bits of its own destination-row range."""
        dev = shard.col.device
        mine = torch.zeros(N, dtype=torch.uint8, device=dev)
        mine[shard.col.long()] = 1
        allb = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allb, mine, group=group)
        mask = torch.zeros(N, dtype=torch.uint8, device=dev)
        slot = 0
        for p in range(self.world):
            if p == self.rank:
                continue
            mask |= allb[p] << slot
            slot += 1
        self.need = mask
        rb, re = shard.row_begin, shard.row_end
        sent = sum(int(((mask[rb:re] >> s_) & 1).sum()) for s_ in range(self.world - 1))
        self.need_fraction = sent / max(1, (re - rb) * (self.world - 1))
        return mask

    def barrier(self, i, group=None):
        """Cross-rank barrier on the current stream: every rank's layer kernel (and its P2P stores) is done."""
        if self.hdls is not None:
            self.hdls[i].barrier()
        else:
            dist.all_reduce(self._tok, group=group)


def exchange_rows(buf, cuts, rank, world, mode="broadcast", group=None):
    """After a layer: every rank has written buf[cuts[rank]:cuts[rank+1]]; make buf complete everywhere."""
    if world == 1:
        return
    if mode == "allreduce":
        rb, re = int(cuts[rank]), int(cuts[rank + 1])
        buf[:rb].zero_(); buf[re:].zero_()
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        return
    sizes = np.diff(np.asarray(cuts))
    if mode == "allgather" and (sizes == sizes[0]).all():
        rb, re = int(cuts[rank]), int(cuts[rank + 1])
        dist.all_gather_into_tensor(buf, buf[rb:re], group=group)       # in place: one collective per layer
        return
    works = []
    for g in range(world):
        rb, re = int(cuts[g]), int(cuts[g + 1])
        if re > rb:
            works.append(dist.broadcast(buf[rb:re], src=g, group=group, async_op=True))
    for w in works:
        w.wait()


def sharded_forward(layer_fn, num_layers, x, shard: Shard, hidden, bufs=None, exchange="broadcast", group=None,
                    head_fn=None):
    """layer_fn(l, h_in, out, shard) must write out[shard.row_begin:shard.row_end] for layer l.
    Returns (h_full, head_fn(h_full, shard) or None)."""
    N = x.shape[0]
    if bufs is None:
        bufs = [torch.empty(N, hidden, dtype=x.dtype, device=x.device) for _ in range(min(2, num_layers))]
    h = x
    for l in range(num_layers):
        out = bufs[l % len(bufs)]
        layer_fn(l, h, out, shard)
        exchange_rows(out, shard.cuts, shard.rank, shard.world, exchange, group)
        h = out
    return h, (head_fn(h, shard) if head_fn is not None else None)


def root_parallel_search(search_fn, rank, world, group=None):
    """search_fn(seed_offset) -> (root_n int32 [A], root_w fp32 [A]).  Returns merged (n, w) on every rank;
    the fp32 sums are accumulated in rank order so every rank gets bit-identical statistics."""
    n, w = search_fn(rank)
    n = torch.as_tensor(np.asarray(n)).clone(); w = torch.as_tensor(np.asarray(w)).clone()
    if world == 1:
        return n.numpy(), w.numpy()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    ns = [torch.empty_like(n, device=dev) for _ in range(world)]
    ws = [torch.empty_like(w, device=dev) for _ in range(world)]
    dist.all_gather(ns, n.to(dev), group=group)
    dist.all_gather(ws, w.to(dev), group=group)
    tot_n = torch.zeros_like(ns[0]); tot_w = torch.zeros_like(ws[0])
    for g in range(world):
        tot_n += ns[g]
        tot_w = tot_w + ws[g]
    return tot_n.cpu().numpy(), tot_w.cpu().numpy()


# ----------------------------------------------------------------------------------------------- GPU side
def gpu_synthetic_graph(N, E, seed, device, relabel=False, family="hub_src"):
    """relabel=True applies a random vertex relabeling (an isomorphic graph): the hub nodes, which every shard
    references, are then spread evenly over the row blocks instead of all living in rank 0's block -- the usual
    partitioning pre-step that balances the per-rank exchange volume.
    Same distribution as graph.synthetic_graph (dst ~ U, src = floor(N u^3), t ~ U[0,60), conf ~ U[.5,1]),
    generated on the GPU so every rank can build the N x 10M-edge graph in about a second.  The CUDA Philox
    generator is deterministic per (seed, call order), so all ranks hold the same graph.
    family: "hub_src" (the headline generator), "hub_dst" (roles swapped: rows with 10^4..10^5 in-edges, SURVEY.md 8d's
    stress variant) or "uniform" (src and dst uniform: no hub rows for the L2 to hold)."""
    gen = torch.Generator(device=device).manual_seed(seed)
    uni = torch.randint(0, N, (E,), generator=gen, device=device, dtype=torch.int64)
    skew = (N * torch.rand(E, generator=gen, device=device, dtype=torch.float64) ** 3).long().clamp_(max=N - 1)
    if family == "hub_src":
        dst, src = uni, skew
    elif family == "hub_dst":
        dst, src = skew, uni
    elif family == "uniform":
        dst, src = uni, torch.randint(0, N, (E,), generator=gen, device=device, dtype=torch.int64)
    else:
        raise ValueError(f"unknown graph family {family!r}")
    del uni, skew
    t = torch.rand(E, generator=gen, device=device) * G.WINDOW
    conf = 0.5 + 0.5 * torch.rand(E, generator=gen, device=device)
    if relabel:
        perm = torch.randperm(N, generator=gen, device=device)
        src, dst = perm[src], perm[dst]
    order = torch.argsort(t, stable=True)
    order = order[torch.argsort(dst[order], stable=True)]
    src, dst, t, conf = src[order], dst[order], t[order], conf[order]
    counts = torch.bincount(dst, minlength=N)
    rowptr = torch.zeros(N + 1, dtype=torch.int32 if E < 2 ** 31 else torch.int64, device=device)
    rowptr[1:] = torch.cumsum(counts, 0)
    ew = conf * torch.exp(-(G.WINDOW - t) / G.TAU)
    x = torch.randn(N, G.F_IN, generator=torch.Generator(device=device).manual_seed(seed + 1), device=device)
    return rowptr, src.to(torch.int32), ew.float().contiguous(), x


def cuda_layer_fn(model):
    """layer_fn for sharded_forward backed by the CUDA kernels (edge block addressed via edge_base)."""
    def fn(l, h, out, shard, score_out=None, peer_outs=None, multicast_ptr=0, peer_need=None):
        model.layer_forward(l, h, shard.rowptr, shard.col, shard.ew, out=out, row_begin=shard.row_begin,
                            row_end=shard.row_end, edge_base=shard.edge_base, score_out=score_out,
                            reuse_long_scan=l > 0, peer_outs=peer_outs, multicast_ptr=multicast_ptr,
                            peer_need=peer_need)
    return fn


def bench_sharded(model, args, world, rank, local_rank, dev, workload_config, algorithmic_bytes_layer, measured_peaks,
                  ClockSampler, physical_gpu_index, run_mcts_bench):
    """bench.py's N > 1 arm (weak scaling: N x (1M nodes, 10M edges), one exchange per layer)."""
    from bench import N_NODES, N_EDGES, HIDDEN, LAYERS
    N, E = N_NODES * world, N_EDGES * world
    rowptr, col, ew, x = gpu_synthetic_graph(N, E, 20250115, dev, relabel=True)
    shard = Shard(rowptr, col, ew, rank, world, device=dev, cut="rows" if args.exchange == "allgather" else "edges")
    shard_src_col, shard_src_ew = col, ew           # kept only until the exchange mode is settled (fallback re-shards)
    pb, need, exchange_note = None, None, ""
    if args.exchange in ("p2p", "p2p-all", "multicast"):
        try:
            pb = PeerBuffers(N, HIDDEN, dev, rank, world, multicast=args.exchange == "multicast")
            need = pb.build_need_mask(shard, N) if args.exchange == "p2p" else None
            ok = torch.ones(1, device=dev)
        except Exception as e:                      # no peer mapping on this system: fall back to the NCCL exchange
            pb, need, ok = None, None, torch.zeros(1, device=dev)
            exchange_note = f" (peer mapping failed: {type(e).__name__}; fell back to allgather)"
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)   # all ranks take the same path
        if float(ok) == 0.0:
            pb, need = None, None
            args.exchange = "allgather"
            shard = Shard(rowptr, shard_src_col, shard_src_ew, rank, world, device=dev, cut="rows")
    del shard_src_col, shard_src_ew, col, ew
    torch.cuda.empty_cache()
    bufs = pb.bufs if pb else [torch.empty(N, HIDDEN, device=dev) for _ in range(2)]
    score = torch.empty(N, device=dev)
    layer = cuda_layer_fn(model)
    K, W = args.steps, max(args.warmup, 3)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2 * LAYERS + 1)] for _ in range(K)]

    def step(i=None):
        h = x
        if i is not None: ev[i][0].record()
        for l in range(LAYERS):
            out = bufs[l & 1]
            last = l == LAYERS - 1
            fused = pb is not None and not last
            layer(l, h, out, shard, score_out=score if last else None, peer_outs=pb.peers[l & 1] if fused else None,
                  multicast_ptr=pb.mc[l & 1] if fused else 0, peer_need=need if fused else None)
            if i is not None: ev[i][2 * l + 1].record()
            if fused:                                       # rows already written into every peer's buffer by the epilogue
                pb.barrier(l & 1)
            elif not last:                                  # the last layer's rows / scores stay sharded
                exchange_rows(out, shard.cuts, rank, world, args.exchange)
            if i is not None: ev[i][2 * l + 2].record()
            h = out

    for _ in range(W):
        step()
    torch.cuda.synchronize(); dist.barrier()
    sampler = ClockSampler(physical_gpu_index(local_rank)); sampler.start()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0.record()
    for i in range(K):
        step(i)
    t1.record()
    torch.cuda.synchronize(); dist.barrier()
    clocks = sampler.result()
    ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms) / K
    seg = np.array([[ev[i][j].elapsed_time(ev[i][j + 1]) for j in range(2 * LAYERS)] for i in range(K)]).mean(0)
    comp_ms, comm_ms = seg[0::2], seg[1::2]
    seg_all = [torch.zeros(2 * LAYERS, device=dev) for _ in range(world)]
    dist.all_gather(seg_all, torch.tensor(seg, device=dev, dtype=torch.float32))
    per_rank = [[round(float(v), 4) for v in t_.cpu()] for t_ in seg_all]
    peak, peak_src = measured_peaks()
    e_loc, r_loc = shard.edge_end - shard.edge_base, shard.row_end - shard.row_begin
    dom_bytes = algorithmic_bytes_layer(e_loc, r_loc, HIDDEN)
    dom_ms = float(comp_ms[1])
    mcts_local = run_mcts_bench(dev, args, seed=rank)
    roll = torch.tensor([mcts_local["value"]], device=dev)
    dist.all_reduce(roll, op=dist.ReduceOp.SUM)
    mcts_local.update({"value": float(roll), "note": "root-parallel: sum over ranks of independent trees (seed = rank), no collective"})
    # ---- e2e: the same sharded step with this rank's inputs coming from pinned host memory every step
    hx, hrp, hcol, hew = (t.cpu().pin_memory() for t in (x, shard.rowptr, shard.col, shard.ew))
    hscore = torch.empty(r_loc).pin_memory()
    def e2e_step():
        x.copy_(hx, non_blocking=True); shard.rowptr.copy_(hrp, non_blocking=True)
        shard.col.copy_(hcol, non_blocking=True); shard.ew.copy_(hew, non_blocking=True)
        step()
        hscore.copy_(score[shard.row_begin:shard.row_end], non_blocking=True)
        torch.cuda.synchronize()
    e2e_step(); dist.barrier()
    te = time.perf_counter()
    for _ in range(K):
        e2e_step()
    dist.barrier()
    e2e_s = torch.tensor([(time.perf_counter() - te) / K], device=dev)
    dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    h2d = int(hx.numel() * 4 + hrp.numel() * hrp.element_size() + hcol.numel() * 4 + hew.numel() * 4)
    e2e = {"value": E / float(e2e_s), "unit": "edges/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(r_loc * 4),
           "ms_per_step": float(e2e_s) * 1e3, "api": "nerrf_b200.dist sharded step, per-rank pinned inputs (bytes are per rank)"}
    tp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
    traffic = json.load(open(tp)).get("sage_layer_F128_dram_bytes_per_launch") if os.path.exists(tp) else None
    return {"metric": "graphsage_t_edges_per_sec", "value": E / (ms_per_step * 1e-3), "unit": "edges/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (torch CUDA generator, same distribution as the N=1 graph, random vertex relabeling for shard balance)",
            "config": dict(workload_config(world), exchange=args.exchange + exchange_note + (f" [{pb.kind}]" if pb else "") +
                           (f", rows sent to a peer only if it references them ({100 * pb.need_fraction:.0f}% of row x peer pairs)" if need is not None else "")),
            "roofline": {"bound": "hbm", "kernel": "fused GraphSAGE-T layer F=128 (rank 0's edge block)", "achieved": dom_bytes / (dom_ms * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": dom_bytes / (dom_ms * 1e-3) / 1e9 / peak, "traffic": traffic if world == 1 else None,
                         "peak_source": peak_src, "per_layer_compute_ms": [float(v) for v in comp_ms],
                         "per_layer_exchange_ms": [float(v) for v in comm_ms],
                         "per_rank_segments_ms": per_rank,
                         "exchange_bytes_per_layer": int(N * HIDDEN * 4)},
            "cpu_baseline": None,
            "e2e": e2e,
            "gpu_launches": K * LAYERS, "clocks": clocks, "mcts": mcts_local, "algo": args.algo}
