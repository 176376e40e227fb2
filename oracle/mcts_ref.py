"""planner.mcts.search oracle (numpy, bit-exact spec; test infra -- see oracle/__init__.py).

Spec source in the reference (prose only; ai/planner/mcts.py is named in README.md:74 and
ROADMAP.md:84 but does not exist):
  * "MCTS ... 500-1000 simulations", "Timeout: 5 min", input "Graph + anomaly scores +
    predictions", output "Undo plan (file reversions, process kills)"
                                                docs/content/docs/architecture.mdx:62-72
  * undo candidates with cost / confidence     docs/content/docs/threat-model.mdx:205-223
  * reward                                      README.md:115  (see rewards_ref.py)

Frozen spec v0 (ours where the reference is silent; DESIGN.md section "MCTS spec").
UCT (Kocsis & Szepesvari 2006) with LEAF PARALLELISM (Chaslot et al. 2008): every
iteration evaluates ONE leaf with R parallel random rollouts.

State      bitset over A_pad actions (applied = 1; padding bits are 1).  legal = 0 bits.
Tree       node 0 = root.  Per node: visits[node] (iterations through it, int32) and three
           arrays over actions: child_n (rollouts, int32), child_w (sum of rollout values,
           fp32), child_id (int32, -1 = none).
Iteration t = 0..T-1
  select   node = root; while visits[node] > 0 and depth < D and n_legal(state) > 0:
             a* = argmax over legal a of key(a), ties -> lowest a, where
                key(a) = +inf                                   if child_n[node][a] == 0
                       = Q + c * sqrt( lnN[visits[node]] / n )  otherwise,
                Q = child_w / n (fp32 IEEE div), n = float(child_n), lnN[k] = fp32(ln(k*R)),
                each op individually rounded (mul, add, div, sqrt correctly rounded; no FMA)
             push (node, a*); state |= a*; depth += 1
             if child_id[node][a*] == -1: create node (id = number of nodes so far); stop
             node = child_id[node][a*]
  rollouts leaf state s0 at depth d, L0 = n_legal(s0).  For r = 0..R-1:
             s = s0; left = D - d
             if left > 0 and L0 > 0: apply the (r mod L0)-th legal action of s0 (stratified
                first move; ascending a); left -= 1
             for k = 0..left-1: L = n_legal(s); stop if 0;
                x = Philox4x32-10(ctr = (r, k >> 2, t, 0), key = (seed_lo, seed_hi))[k & 3]
                j = (x * L) >> 32;  apply the j-th legal action (ascending a)
             val[r] = (score(s) - lo) * inv_range        (fp32, two rounded ops)
  backup   total = adjacent-pairs binary-tree sum of val[0..R)            (R is a power of 2)
           leaf children (only when the stratified first move happened): for rank q <
             min(R, L0), a = q-th legal action of s0: child_n[leaf][a] += #{r = q mod L0};
             child_w[leaf][a] += val[q] + val[q+L0] + ... added one by one in ascending r
           visits[leaf] += 1
           every (node, a) on the path: child_n += R; child_w += total; visits[node] += 1
Result     root child statistics, and `best` = argmax over root children with n > 0 of
           (n, Q, -a) lexicographic (robust child; Q breaks ties; then lowest index).
"""
import numpy as np

from .philox import philox4x32_10
from . import rewards_ref as RW

F32 = np.float32


def _popcount32(x):
    x = x.astype(np.uint32)
    x = x - ((x >> np.uint32(1)) & np.uint32(0x55555555))
    x = (x & np.uint32(0x33333333)) + ((x >> np.uint32(2)) & np.uint32(0x33333333))
    x = (x + (x >> np.uint32(4))) & np.uint32(0x0F0F0F0F)
    return ((x * np.uint32(0x01010101)) >> np.uint32(24)).astype(np.int64)


def _select_kth_zero(states, k):
    """states uint32 [R, nw]; k int64 [R] (0-based rank among ZERO bits, ascending a).
    Returns action index int64 [R]."""
    R, nw = states.shape
    zeros = 32 - _popcount32(states)                       # [R, nw]
    cum = np.cumsum(zeros, axis=1)                          # inclusive
    word = (cum <= k[:, None]).sum(axis=1)                  # first word with cum > k
    before = np.where(word > 0, np.take_along_axis(cum, np.maximum(word - 1, 0)[:, None], 1)[:, 0], 0)
    rank = k - before                                       # rank within word
    w = ~np.take_along_axis(states, word[:, None], 1)[:, 0]  # set bits = legal
    # find the rank-th set bit of w (binary search on prefix popcounts)
    pos = np.zeros(R, np.int64)
    for width in (16, 8, 4, 2, 1):
        mask = ((np.uint64(1) << (pos + width).astype(np.uint64)) - np.uint64(1)) ^ \
               ((np.uint64(1) << pos.astype(np.uint64)) - np.uint64(1))
        cnt = _popcount32((w.astype(np.uint64) & mask).astype(np.uint32))
        go = cnt <= rank
        rank = np.where(go, rank - cnt, rank)
        pos = np.where(go, pos + width, pos)
    return word * 32 + pos


def _set_bit(states, a, active=None):
    R = states.shape[0]
    rows = np.arange(R)
    word = a >> 5
    bit = (np.uint32(1) << (a & 31).astype(np.uint32)).astype(np.uint32)
    if active is not None:
        bit = np.where(active, bit, np.uint32(0)).astype(np.uint32)
        word = np.where(active, word, 0)
    states[rows, word] |= bit


def tree_sum(val):
    v = np.asarray(val, dtype=F32).copy()
    assert v.size & (v.size - 1) == 0
    while v.size > 1:
        v = (v[0::2] + v[1::2]).astype(F32)
    return v[0]


def ln_table(T, R):
    k = np.arange(T + 2, dtype=np.float64)
    k[0] = 1.0
    out = np.log(k * float(R)).astype(F32)
    out[0] = F32(0.0)
    return out


class Tree:
    def __init__(self, T, A_pad):
        self.visits = np.zeros(T + 1, np.int32)
        self.child_n = np.zeros((T + 1, A_pad), np.int32)
        self.child_w = np.zeros((T + 1, A_pad), F32)
        self.child_id = np.full((T + 1, A_pad), -1, np.int32)
        self.num_nodes = 1


def rollouts(s0, depth, t, R, D, seed, p, size, cost, lo, inv_range, guard=None):
    """One leaf evaluation.  Returns (val fp32 [R], first int64 [R] (-1 if none), L0, final states)."""
    A = np.asarray(p).shape[0]
    NW, chunk, A_pad, nw = RW.layout(A)
    s = np.repeat(np.asarray(s0, np.uint32)[None, :], R, axis=0)
    L0 = int((32 - _popcount32(s[0])).sum())
    left = D - depth
    first = np.full(R, -1, np.int64)
    r = np.arange(R, dtype=np.int64)
    if left > 0 and L0 > 0:
        first = _select_kth_zero(s, r % L0)
        _set_bit(s, first)
        left -= 1
    k0 = np.uint32(seed & 0xFFFFFFFF); k1 = np.uint32((seed >> 32) & 0xFFFFFFFF)
    rnd = None
    for k in range(max(left, 0)):
        L = (32 - _popcount32(s)).sum(axis=1)
        active = L > 0
        if not active.any():
            break
        if (k & 3) == 0:
            rnd = philox4x32_10(r.astype(np.uint32), np.uint32(k >> 2), np.uint32(t), np.uint32(0), k0, k1)
        x = rnd[k & 3].astype(np.uint64)
        j = ((x * L.astype(np.uint64)) >> np.uint64(32)).astype(np.int64)
        j = np.where(active, j, 0)
        # inactive rows (no legal action) keep their state: select on a dummy but mask the set
        a = _select_kth_zero(np.where(active[:, None], s, np.uint32(0)), j)
        _set_bit(s, a, active)
    sc = RW.score(s, p, size, cost, guard)
    val = ((sc - F32(lo)).astype(F32) * F32(inv_range)).astype(F32)
    return val, first, L0, s


def search(p, size, cost, R=4096, D=50, T=64, seed=0, c=np.sqrt(2.0), root_state=None,
           bounds=None, return_tree=False, guard=None):
    p = np.asarray(p, F32); size = np.asarray(size, F32); cost = np.asarray(cost, F32)
    A = p.shape[0]
    NW, chunk, A_pad, nw = RW.layout(A)
    assert R >= 1 and (R & (R - 1)) == 0, "R must be a power of two"
    root = RW.empty_state(A) if root_state is None else (np.asarray(root_state, np.uint32) | RW.empty_state(A))
    lo, inv = RW.reward_bounds(p, size, cost, root, guard) if bounds is None else bounds
    lnN = ln_table(T, R)
    c32 = F32(c)
    tree = Tree(T, A_pad)
    legal_idx_cache = np.arange(A_pad)
    for t in range(T):
        node = 0; state = root.copy(); depth = 0; path = []
        created = False
        while True:
            if tree.visits[node] == 0:
                break
            legal = ~RW.unpack_bits(state[None, :])[0]
            if depth >= D or not legal.any():
                break
            n = tree.child_n[node]
            nf = n.astype(F32)
            with np.errstate(divide="ignore", invalid="ignore"):
                q = (tree.child_w[node] / nf).astype(F32)
                ratio = (lnN[tree.visits[node]] / nf).astype(F32)
                key = (q + (c32 * np.sqrt(ratio).astype(F32)).astype(F32)).astype(F32)
            key = np.where(n == 0, F32(np.inf), key)
            key = np.where(legal, key, -np.inf)
            a = int(np.argmax(key))          # first maximal index == lowest a on ties
            path.append((node, a))
            state[a >> 5] |= np.uint32(1) << np.uint32(a & 31)
            depth += 1
            if tree.child_id[node][a] == -1:
                nid = tree.num_nodes
                tree.num_nodes += 1
                tree.child_id[node][a] = nid
                node = nid
                created = True
                break
            node = int(tree.child_id[node][a])
        leaf = node
        val, first, L0, _ = rollouts(state, depth, t, R, D, seed, p, size, cost, lo, inv, guard)
        total = tree_sum(val)
        if first[0] >= 0:
            nq = min(R, L0)
            for q_ in range(nq):
                a = int(first[q_])
                idx = np.arange(q_, R, L0)
                w = tree.child_w[leaf][a]
                for i in idx:
                    w = F32(w + val[i])
                tree.child_w[leaf][a] = w
                tree.child_n[leaf][a] += len(idx)
        tree.visits[leaf] += 1
        for (nd, a) in path:
            tree.child_n[nd][a] += R
            tree.child_w[nd][a] = F32(tree.child_w[nd][a] + total)
            tree.visits[nd] += 1
    root_n = tree.child_n[0][:A].copy()
    root_w = tree.child_w[0][:A].copy()
    best = best_child(root_n, root_w)
    out = {"root_n": root_n, "root_w": root_w, "best": best, "num_nodes": tree.num_nodes,
           "lo": lo, "inv_range": inv}
    if return_tree:
        out["tree"] = tree
    return out


def best_child(root_n, root_w):
    """argmax of (n, Q, -a) over children with n > 0; -1 if none."""
    n = np.asarray(root_n); w = np.asarray(root_w, F32)
    best = -1; bn = 0; bq = F32(0)
    for a in range(n.shape[0]):
        if n[a] <= 0:
            continue
        q = F32(w[a] / F32(n[a]))
        if best < 0 or n[a] > bn or (n[a] == bn and q > bq):
            best, bn, bq = a, int(n[a]), q
    return best
