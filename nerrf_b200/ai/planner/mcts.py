"""ai.planner.mcts -- MCTS rollback planner (CUDA persistent leaf-parallel kernel, sm_100a).

Reference surface: ai/planner/mcts.py `search` (README.md:74,115; ROADMAP.md:84 -- named, never
written).  Behaviour: "500-1000 simulations", "Timeout: 5 min", input "Graph + anomaly scores +
predictions", output "Undo plan (file reversions, process kills)"
(docs/content/docs/architecture.mdx:62-72); candidates threat-model.mdx:205-223.
Frozen spec: DESIGN.md "MCTS spec" (leaf-parallel UCT, R rollouts per iteration, Philox RNG,
bit-exact integer statistics).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np
import torch

from ... import _lib as L
from . import rewards as RW
from .rewards import Actions


@dataclass
class SearchResult:
    root_n: np.ndarray      # int32 [A]  rollouts through each root child
    root_w: np.ndarray      # fp32  [A]  sum of rollout values
    best: int               # recommended action (-1 if none)
    num_nodes: int
    lo: float
    inv_range: float
    rollouts: int

    @property
    def root_q(self):
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.where(self.root_n > 0, self.root_w / self.root_n.astype(np.float32), np.float32(0)).astype(np.float32)


def ln_table(T, R) -> np.ndarray:
    k = np.arange(T + 2, dtype=np.float64)
    k[0] = 1.0
    out = np.log(k * float(R)).astype(np.float32)
    out[0] = np.float32(0.0)
    return out


def best_child(root_n, root_w) -> int:
    """argmax over children with n > 0 of (n, Q, -a): robust child, Q breaks ties, then lowest index."""
    n = np.asarray(root_n); w = np.asarray(root_w, np.float32)
    best, bn, bq = -1, 0, np.float32(0)
    for a in np.nonzero(n > 0)[0]:
        q = np.float32(w[a] / np.float32(n[a]))
        if best < 0 or n[a] > bn or (n[a] == bn and q > bq):
            best, bn, bq = int(a), int(n[a]), q
    return best


def merge_root_stats(results):
    """Root parallelism: sum the per-tree root statistics in rank order."""
    n = np.zeros_like(results[0].root_n)
    w = np.zeros_like(results[0].root_w)
    for r in results:
        n = n + r.root_n
        w = (w + r.root_w).astype(np.float32)
    return n, w, best_child(n, w)


def search(actions: Actions, scorer=None, n_rollouts: int = 4096, depth: int = 50, seed: int = 0,
           c: float = math.sqrt(2.0), iterations: int = 64, root_state=None, device=None, host_call: bool = False
           ) -> SearchResult:
    """One tree search.  `scorer` must be None or ai.planner.rewards.score: the reward is evaluated
    inside the kernel (batched, R states per iteration).  host_call=True goes through the
    host-buffer C-ABI entry (copies inside the call)."""
    if scorer is not None and scorer is not RW.score:
        raise NotImplementedError("the CUDA planner evaluates ai.planner.rewards.score in-kernel; "
                                  "custom scorers are not supported")
    A = actions.A
    NW, _, A_pad, nw = RW.layout(A)
    R, D, T = int(n_rollouts), int(depth), int(iterations)
    if R < 1 or (R & (R - 1)):
        raise ValueError("n_rollouts must be a power of two")
    root = RW.empty_state(A) if root_state is None else (np.asarray(root_state, np.uint32) | RW.empty_state(A))
    lo, inv = RW.reward_bounds(actions, root)
    lnN = ln_table(T, R)
    lib = L.lib()
    if host_call:
        root_n = np.zeros(A_pad, np.int32); root_w = np.zeros(A_pad, np.float32); nn = np.zeros(1, np.int32)
        as_p = lambda a: a.ctypes.data_as(C.c_void_p)
        L.check(lib.nerrf_mcts_search_host(as_p(actions.p), as_p(actions.size), as_p(actions.cost), A, as_p(root), R, D, T,
                                           C.c_uint64(seed), float(c), float(lo), float(inv), as_p(lnN), as_p(root_n),
                                           as_p(root_w), as_p(nn)), "nerrf_mcts_search_host")
        return SearchResult(root_n[:A].copy(), root_w[:A].copy(), best_child(root_n[:A], root_w[:A]), int(nn[0]),
                            float(lo), float(inv), R * T)
    device = torch.device(device or "cuda")
    if device.type != "cuda":
        raise L.NerrfError("mcts.search needs a CUDA device (no CPU fallback)")
    with torch.cuda.device(device):
        p, size, cost = actions.device_arrays(device)
        d_root = torch.from_numpy(root.view(np.int32)).to(device)
        d_ln = torch.from_numpy(lnN).to(device)
        d_n = torch.empty(A_pad, device=device, dtype=torch.int32)
        d_w = torch.empty(A_pad, device=device, dtype=torch.float32)
        d_nn = torch.empty(1, device=device, dtype=torch.int32)
        need = C.c_size_t()
        L.check(lib.nerrf_mcts_workspace_bytes(A, T, R, C.byref(need)), "nerrf_mcts_workspace_bytes")
        ws = torch.empty(need.value, device=device, dtype=torch.uint8)
        L.check(lib.nerrf_mcts_search(L.ptr(p), L.ptr(size), L.ptr(cost), A, L.ptr(d_root), R, D, T, C.c_uint64(seed),
                                      float(c), float(lo), float(inv), L.ptr(d_ln), L.ptr(d_n), L.ptr(d_w), L.ptr(d_nn),
                                      L.ptr(ws), need.value, L.current_stream_ptr()), "nerrf_mcts_search")
        root_n = d_n.cpu().numpy()[:A].copy(); root_w = d_w.cpu().numpy()[:A].copy(); nn = int(d_nn.cpu())
    return SearchResult(root_n, root_w, best_child(root_n, root_w), nn, float(lo), float(inv), R * T)


@dataclass
class Plan:
    actions: list           # committed action indices, in order
    scores: list            # exact rewards.score after each commit (scores[0] = initial state)
    searches: list          # SearchResult per step


def ranked_children(root_n, root_w):
    """Root children with n > 0 ordered by (n desc, Q desc, index asc) -- best_child() is element 0."""
    n = np.asarray(root_n); w = np.asarray(root_w, np.float32)
    idx = np.nonzero(n > 0)[0]
    q = (w[idx] / n[idx].astype(np.float32)).astype(np.float32)
    order = np.lexsort((idx, -q.astype(np.float64), -n[idx].astype(np.int64)))
    return idx[order]


def plan(actions: Actions, max_steps: int | None = None, n_rollouts: int = 4096, depth: int = 50, seed: int = 0,
         c: float = math.sqrt(2.0), iterations: int = 64, device=None) -> Plan:
    """Undo plan = repeated search / commit / re-root.  Each step the search ranks the root
    children; the candidates are then VALIDATED with the exact reward (one batched
    rewards.score call over all candidate next-states, mirroring the reference's "sandbox
    validates, then apply" gate, architecture.mdx:81-86): the highest-ranked candidate whose
    exact reward improves on the current state is committed; the plan ends when none does."""
    A = actions.A
    state = RW.empty_state(A)
    max_steps = depth if max_steps is None else max_steps
    cur = float(RW.score(state[None, :], actions, device=device).cpu()[0])
    out = Plan([], [cur], [])
    for step in range(max_steps):
        res = search(actions, None, n_rollouts, max(depth - step, 1), seed + step, c, iterations, state, device)
        out.searches.append(res)
        cand = ranked_children(res.root_n, res.root_w)
        if cand.size == 0:
            break
        nxt = np.repeat(state[None, :], cand.size, axis=0)
        nxt[np.arange(cand.size), cand >> 5] |= (np.uint32(1) << (cand & 31).astype(np.uint32)).astype(np.uint32)
        sc = RW.score(nxt, actions, device=device).cpu().numpy()
        better = np.nonzero(sc > np.float32(cur))[0]
        if better.size == 0:
            break
        k = int(better[0])
        state, cur = nxt[k].copy(), float(sc[k])
        out.actions.append(int(cand[k])); out.scores.append(cur)
    return out
