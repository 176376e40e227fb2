"""World-size-2 (and 3) gloo tests of the sharded GraphSAGE-T forward and root-parallel MCTS merge.
The compute is the oracle (CPU); what is under test is the host-side sharding / exchange logic."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nerrf_b200 import graph as G
from nerrf_b200 import dist as ND
from oracle import sage_ref as S
from oracle import mcts_ref as M


def _worker(rank, world, port, exchange, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = G.synthetic_graph(N=3000, E=30000, seed=5, hub="dst")          # skewed rows: uneven cuts
        P = S.make_params(32, 128, 3, seed=1)
        t = lambda a: torch.from_numpy(a)
        shard = ND.Shard(g.rowptr, g.col, g.ew, rank, world, cut="rows" if exchange == "allgather" else "edges")
        full_col, full_ew = t(g.col), t(g.ew)
        assert torch.equal(shard.col, full_col[shard.edge_base:shard.edge_end])

        def layer_fn(l, h, out, sh):
            W, b = P["layers"][l]
            out[sh.row_begin:sh.row_end] = S.layer(h, t(g.rowptr), full_col, full_ew, W, b,
                                                    row_begin=sh.row_begin, row_end=sh.row_end)

        def head_fn(h, sh):
            return torch.sigmoid(h[sh.row_begin:sh.row_end] @ P["node_w"] + P["node_b"])

        h, score = ND.sharded_forward(layer_fn, 3, t(g.x), shard, 128, exchange=exchange, head_fn=head_fn)
        hw, scw = S.forward(P, t(g.x), t(g.rowptr), full_col, full_ew)
        ok = torch.allclose(h, hw, rtol=1e-5, atol=1e-6) and torch.allclose(score, scw[shard.row_begin:shard.row_end], atol=1e-6)
        # need mask of the fused exchange: bit (peer slot) set <=> that peer's edge block references the row
        pb = ND.PeerBuffers.__new__(ND.PeerBuffers); pb.rank, pb.world = rank, world
        mask = pb.build_need_mask(shard, g.num_nodes).numpy()
        cuts = shard.cuts
        slot = 0
        for p_ in range(world):
            if p_ == rank:
                continue
            e0, e1 = int(g.rowptr[cuts[p_]]), int(g.rowptr[cuts[p_ + 1]])
            want_bits = np.zeros(g.num_nodes, bool); want_bits[g.col[e0:e1]] = True
            ok = ok and np.array_equal(((mask >> slot) & 1).astype(bool), want_bits)
            slot += 1
        ok = ok and 0.0 < pb.need_fraction <= 1.0
        # root-parallel MCTS merge: every rank ends with the same merged statistics
        rng = np.random.default_rng(0)
        p = rng.beta(0.5, 0.5, 40).astype(np.float32); size = rng.lognormal(0.7, 1, 40).astype(np.float32)
        cost = np.ones(40, np.float32)
        fn = lambda off: (lambda r: (r["root_n"], r["root_w"]))(M.search(p, size, cost, R=64, D=10, T=6, seed=7 + off))
        n, w = ND.root_parallel_search(fn, rank, world)
        want_n = sum(M.search(p, size, cost, R=64, D=10, T=6, seed=7 + g_)["root_n"] for g_ in range(world))
        ok = ok and np.array_equal(n, want_n)
        q.put((rank, bool(ok), shard.row_begin, shard.row_end, w.tobytes()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,exchange", [(2, "broadcast"), (2, "allreduce"), (3, "broadcast"), (2, "allgather"), (3, "allgather")])
def test_sharded_forward_matches_single_process(world, exchange):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world * 7 + {"allreduce": 3, "broadcast": 0, "allgather": 5}[exchange]
    procs = [ctx.Process(target=_worker, args=(r, world, port, exchange, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    res.sort()
    assert all(r[1] for r in res), res
    assert res[0][2] == 0 and res[-1][3] == 3000 and all(a[3] == b[2] for a, b in zip(res, res[1:]))
    assert len({r[4] for r in res}) == 1            # bit-identical merged MCTS value sums on every rank


def _worker_pipeline_pieces(rank, world, port, q):
    """Host-side logic of the multi-GPU pipeline (pipeline.DistContext) under gloo: the LSTM batch split + all-gather
    returns exactly the unsplit result in the original order; component-aligned cuts never split a component."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nerrf_b200 import pipeline
        ctx = pipeline.DistContext(rank, world)
        rng = np.random.default_rng(3)
        seq = rng.standard_normal((37, 5, 4)).astype(np.float32); ln = rng.integers(0, 6, 37).astype(np.int32)
        model = lambda s, l: torch.stack([s.sum((1, 2)), l.float()], 1)          # any per-row function
        got = pipeline._lstm_probs(model, seq, ln, torch.device("cpu"), ctx)
        want = model(torch.from_numpy(seq), torch.from_numpy(ln))
        ok = torch.equal(got, want)
        tiny = pipeline._lstm_probs(model, seq[:1], ln[:1], torch.device("cpu"), ctx)   # fewer rows than ranks: no split
        ok = ok and torch.equal(tiny, want[:1])
        # component-aligned cuts: 50 components of 64 nodes, uneven edges per component
        S, K = 64, 50
        deg = np.repeat(rng.integers(1, 30, K), S)
        rowptr = np.zeros(S * K + 1, np.int64); np.cumsum(deg, out=rowptr[1:])
        col = np.zeros(int(rowptr[-1]), np.int32); ew = np.ones(int(rowptr[-1]), np.float32)
        sh = ND.Shard(rowptr, col, ew, rank, world, align=S)
        ok = ok and all(int(c) % S == 0 for c in sh.cuts) and sh.cuts[0] == 0 and sh.cuts[-1] == S * K
        ok = ok and all(b >= a for a, b in zip(sh.cuts, sh.cuts[1:]))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_pipeline_batch_split_and_component_aligned_cuts():
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 91
    procs = [ctx.Process(target=_worker_pipeline_pieces, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
