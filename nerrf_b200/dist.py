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


import numpy as np
import torch
import torch.distributed as dist

from . import graph as G


class Shard:
    """What one rank holds: full rowptr, its edge block of col / ew (indexed from edge_base)."""

    def __init__(self, rowptr, col, ew, rank, world, device=None, cut="edges", align=1):
        """cut="edges": edge-balanced row-aligned cuts (skewed graphs); cut="rows": equal row blocks (when the
        in-degree is uniform these are edge-balanced too, and the exchange can be one in-place all-gather).
        align: cuts are rounded to a multiple of `align` rows -- the component size of a trace-structured graph
        (one process + its files, contiguous node ids), so that no component straddles two ranks and the per-layer
        exchange has (almost) nothing to send."""
        rp = rowptr.cpu().numpy() if isinstance(rowptr, torch.Tensor) else np.asarray(rowptr)
        n = rp.shape[0] - 1
        if cut == "rows":
            self.cuts = np.array([(n * g) // world for g in range(world + 1)], dtype=np.int64)
        else:
            self.cuts = G.edge_balanced_row_cuts(rp, world)
        if align > 1:
            c = (np.round(self.cuts / align).astype(np.int64) * align).clip(0, n)
            c[0], c[-1] = 0, n
            self.cuts = np.maximum.accumulate(c)
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


def gpu_trace_graph(n_components, device, seed=7, files=63, ev_lo=3, ev_hi=8):
    """Trace-STRUCTURED synthetic graph: what the graph constructor produces for a fleet of processes, each touching its
    own files (the LockBit traces benchmarks/m{0,1}/results/*_trace.jsonl are one such component: 1 pid + 57 / 98 paths).
    Component c = node c*S (the process) + S-1 file nodes, S = files + 1, node ids contiguous per component; every file
    has ev_lo..ev_hi-1 events, each event one edge process->file and one file->process (nerrf_b200.graph
    graph_from_events), t ~ U[0, 60 s), conf = 1.  -> (rowptr, col, ew, x, S).  Generated on the GPU, deterministic per
    seed, so every rank holds the same graph."""
    S = files + 1
    N = n_components * S
    gen = torch.Generator(device=device).manual_seed(seed)
    n_files = n_components * files
    ev = torch.randint(ev_lo, ev_hi, (n_files,), generator=gen, device=device)
    fidx = torch.repeat_interleave(torch.arange(n_files, device=device), ev)
    file_node = (fidx // files) * S + 1 + fidx % files
    pid_node = (fidx // files) * S
    t1 = torch.rand(fidx.numel(), generator=gen, device=device) * G.WINDOW
    src = torch.cat([pid_node, file_node]); dst = torch.cat([file_node, pid_node]); t = torch.cat([t1, t1])
    del fidx, file_node, pid_node, t1
    order = torch.argsort(t, stable=True)
    order = order[torch.argsort(dst[order], stable=True)]
    src, dst, t = src[order], dst[order], t[order]
    E = src.numel()
    rowptr = torch.zeros(N + 1, dtype=torch.int32, device=device)
    rowptr[1:] = torch.cumsum(torch.bincount(dst, minlength=N), 0)
    ew = torch.exp(-(G.WINDOW - t) / G.TAU).float().contiguous()
    x = torch.randn(N, G.F_IN, generator=torch.Generator(device=device).manual_seed(seed + 1), device=device)
    return rowptr, src.to(torch.int32).contiguous(), ew, x, S


class ShardedSage:
    """The 1-D edge-block sharded GraphSAGE-T forward of one rank: graph shard, peer-mapped
    embedding buffers, need mask, and the per-layer loop with the exchange fused into the layer kernel's epilogue
    (exchange="p2p" / "p2p-all" / "multicast") or done by NCCL ("allgather" / "broadcast" / "allreduce").

    Buffer reuse across steps: the last layer has no exchange, so a fast rank could start step i+1 and store layer-0
    rows into a peer's buffer while that peer still READS it as the input of its step-i last layer (even layer
    counts) -- every step therefore ends with one more cross-rank barrier."""

    def __init__(self, model, rowptr, col, ew, rank, world, device, exchange="p2p", align=1, hidden=128):
        self.model, self.rank, self.world, self.dev, self.hidden = model, rank, world, device, hidden
        self.exchange, self.note = exchange, ""
        self.N = rowptr.numel() - 1
        self.L = model.num_layers
        cut = "rows" if exchange == "allgather" else "edges"
        self.shard = Shard(rowptr, col, ew, rank, world, device=device, cut=cut, align=align)
        self.pb = self.need = None
        if exchange in ("p2p", "p2p-all", "multicast"):
            try:
                self.pb = PeerBuffers(self.N, hidden, device, rank, world, multicast=exchange == "multicast")
                self.need = self.pb.build_need_mask(self.shard, self.N) if exchange == "p2p" else None
                ok = torch.ones(1, device=device)
            except Exception as e:                      # no peer mapping on this system: fall back to the NCCL exchange
                self.pb, self.need, ok = None, None, torch.zeros(1, device=device)
                self.note = f" (peer mapping failed: {type(e).__name__}; fell back to allgather)"
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)   # all ranks take the same path
            if float(ok) == 0.0:
                self.pb = self.need = None
                self.exchange = "allgather"
                self.shard = Shard(rowptr, col, ew, rank, world, device=device, cut="rows")
        self.bufs = self.pb.bufs if self.pb else [torch.empty(self.N, hidden, device=device) for _ in range(2)]
        self.score = torch.empty(self.N, device=device)
        self.layer = cuda_layer_fn(model)
        self.x = None
        self._xsym = None

    # -- inputs ---------------------------------------------------------------------------------------------------
    def set_x(self, x):
        """Resident features: the full [N, F] matrix already on this device."""
        self.x = x

    def sharded_upload(self, hx_own, h_rowptr_own, h_col, h_ew):
        """The e2e input path: this rank copies only ITS rows of x (and its slices of rowptr / col / ew) from pinned host
        memory, then the ranks complete each other's x over NVLink -- a row goes to the peers whose edge block
        references it (the need mask), written straight into their x buffers.  Replaces every rank pulling the whole x
        through its own PCIe link."""
        sh = self.shard
        if self._xsym is None:
            self._init_x_exchange(hx_own.shape[1])
        xs = self._xsym
        xs[sh.row_begin:sh.row_end].copy_(hx_own, non_blocking=True)
        sh.rowptr[sh.row_begin:sh.row_end + 1].copy_(h_rowptr_own, non_blocking=True)
        sh.col.copy_(h_col, non_blocking=True); sh.ew.copy_(h_ew, non_blocking=True)
        if self._x_peers is not None:
            for p_x, idx in zip(self._x_peers, self._x_idx):
                if idx.numel():
                    p_x[idx] = xs[idx]                                  # remote stores over NVLink
            self._x_hdl.barrier()
        else:
            exchange_rows(xs, sh.cuts, self.rank, self.world, "broadcast")
        self.x = xs

    def _init_x_exchange(self, F):
        sh = self.shard
        self._x_peers = None
        if self.pb is not None and self.pb.hdls is not None:
            import torch.distributed._symmetric_memory as symm
            self._xsym = symm.empty(self.N, F, dtype=torch.float32, device=self.dev)
            self._x_hdl = symm.rendezvous(self._xsym, dist.group.WORLD)
            self._x_peers = [self._x_hdl.get_buffer(p, (self.N, F), torch.float32) for p in range(self.world) if p != self.rank]
            own = torch.arange(sh.row_begin, sh.row_end, device=self.dev)
            if self.need is not None:
                m = self.need[sh.row_begin:sh.row_end]
                self._x_idx = [own[((m >> s_) & 1).bool()] for s_ in range(self.world - 1)]
            else:
                self._x_idx = [own for _ in range(self.world - 1)]
        else:
            self._xsym = torch.empty(self.N, F, dtype=torch.float32, device=self.dev)

    # -- one forward ----------------------------------------------------------------------------------------------
    def step(self, ev=None, snapshots=None):
        """One sharded forward; the node scores of the own rows land in self.score[row_begin:row_end].
        ev: list of 2L+1 CUDA events (compute / exchange segments); snapshots: list that receives a clone of every
        layer's full output buffer after its exchange (parity checks only)."""
        pb, sh, L = self.pb, self.shard, self.L
        h = self.x
        if ev is not None: ev[0].record()
        for l in range(L):
            out = self.bufs[l & 1]
            last = l == L - 1
            fused = pb is not None and not last
            self.layer(l, h, out, sh, score_out=self.score if last else None, peer_outs=pb.peers[l & 1] if fused else None,
                       multicast_ptr=pb.mc[l & 1] if fused else 0, peer_need=self.need if fused else None)
            if ev is not None: ev[2 * l + 1].record()
            if fused:                                       # rows already written into every peer's buffer by the epilogue
                pb.barrier(l & 1)
            elif not last:                                  # the last layer's rows / scores stay sharded
                exchange_rows(out, sh.cuts, self.rank, self.world, self.exchange)
            elif pb is not None:
                pb.barrier(l & 1)                           # nobody overwrites a buffer a peer is still reading (next step)
            if ev is not None: ev[2 * l + 2].record()
            if snapshots is not None:
                snapshots.append(out.clone())
            h = out
        return self.score

    def describe(self):
        pb = self.pb
        return (self.exchange + self.note + (f" [{pb.kind}]" if pb else "") +
                (f", rows sent to a peer only if it references them ({100 * pb.need_fraction:.1f}% of row x peer pairs)"
                 if self.need is not None else ""))

    def exchange_bytes_per_layer(self):
        """(egress, ingress) bytes of this rank per exchanged layer."""
        sh, H = self.shard, self.hidden
        rows = sh.row_end - sh.row_begin
        if self.need is not None:
            eg = int(sum(int(((self.need[sh.row_begin:sh.row_end] >> s_) & 1).sum()) for s_ in range(self.world - 1))) * H * 4
            t = torch.tensor([eg], device=self.dev, dtype=torch.float64)
            dist.all_reduce(t)                               # symmetric on average: report the mean as ingress
            return eg, int(float(t) / self.world)
        if self.pb is not None or self.exchange in ("allgather", "broadcast"):
            return rows * (self.world - 1) * H * 4, (self.N - rows) * H * 4
        return self.N * H * 4, self.N * H * 4

    # -- parity ---------------------------------------------------------------------------------------------------
    def parity_vs_single_gpu(self, rowptr, col, ew, x):
        """Sharded forward == single-GPU forward of the same kernel on the SAME full graph, bit for bit: this rank's own
        rows of every layer and of the scores, and -- for every exchanged layer -- every row this rank READS as a source
        (i.e. every row a peer had to deliver).  All arguments are full-graph tensors on this device."""
        model, sh, L = self.model, self.shard, self.L
        snaps = []
        self.set_x(x) if self.x is None else None
        self.step(snapshots=snaps)
        torch.cuda.synchronize()
        reads = torch.zeros(self.N, dtype=torch.bool, device=self.dev)
        reads[sh.col.long()] = True
        reads[sh.row_begin:sh.row_end] = True
        ridx = reads.nonzero().squeeze(1)
        own = slice(sh.row_begin, sh.row_end)
        h = x
        own_ok, read_ok, worst = True, True, 0.0
        bufs = [torch.empty(self.N, self.hidden, device=self.dev) for _ in range(2)]
        score = torch.empty(self.N, device=self.dev)
        for l in range(L):
            ref = bufs[l & 1]
            model.layer_forward(l, h, rowptr, col, ew, out=ref, score_out=score if l == L - 1 else None, reuse_long_scan=l > 0)
            own_ok &= bool(torch.equal(snaps[l][own], ref[own]))
            worst = max(worst, float((snaps[l][own] - ref[own]).abs().max()) if sh.row_end > sh.row_begin else 0.0)
            if l < L - 1:
                read_ok &= bool(torch.equal(snaps[l][ridx], ref[ridx]))
            h = ref
        own_ok &= bool(torch.equal(self.score[own], score[own]))
        flags = torch.tensor([int(own_ok), int(read_ok)], device=self.dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        w = torch.tensor([worst], device=self.dev); dist.all_reduce(w, op=dist.ReduceOp.MAX)
        nread = torch.tensor([float(ridx.numel() - (sh.row_end - sh.row_begin))], device=self.dev); dist.all_reduce(nread)
        return {"own_rows_bit_exact": bool(flags[0]), "read_rows_bit_exact": bool(flags[1]), "layers": L,
                "max_abs_diff_own_rows": float(w), "remote_rows_read_checked_per_layer": int(float(nread)),
                "what": "sharded forward vs the single-GPU forward of the full graph on every rank: own rows of every layer "
                        "+ scores, and every remotely produced row a rank reads as a source"}, (h, score)


def timed_sharded_run(ss: "ShardedSage", K, W, sampler_factory=None):
    """W warm-up + K timed sharded forwards (barrier + synchronize on both sides, CUDA events, max over ranks).
    -> dict(ms_per_step, compute_ms [L], exchange_ms [L], per_rank_segments, clocks)."""
    L, dev, world = ss.L, ss.dev, ss.world
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2 * L + 1)] for _ in range(K)]
    for _ in range(W):
        ss.step()
    torch.cuda.synchronize(); dist.barrier()
    sampler = sampler_factory() if sampler_factory else None
    if sampler: sampler.start()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0.record()
    for i in range(K):
        ss.step(ev[i])
    t1.record()
    torch.cuda.synchronize(); dist.barrier()
    clocks = sampler.result() if sampler else None
    ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    seg = np.array([[ev[i][j].elapsed_time(ev[i][j + 1]) for j in range(2 * L)] for i in range(K)]).mean(0)
    seg_all = [torch.zeros(2 * L, device=dev) for _ in range(world)]
    dist.all_gather(seg_all, torch.tensor(seg, device=dev, dtype=torch.float32))
    return {"ms_per_step": float(ms) / K, "compute_ms": [float(v) for v in seg[0::2]], "exchange_ms": [float(v) for v in seg[1::2]],
            "per_rank_segments_ms": [[round(float(v), 4) for v in t_.cpu()] for t_ in seg_all], "clocks": clocks}
