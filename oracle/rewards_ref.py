"""rewards.score oracle (numpy, bit-exact fp32 spec; test infra -- see oracle/__init__.py).

Spec source in the reference (prose only; no code exists):
  * ``Reward = -(data_loss + 0.1 x downtime)``            README.md:115
  * "restoration gain - side effects"                     docs/content/docs/architecture.mdx:71
  * per-candidate cost / confidence / reward example      docs/content/docs/threat-model.mdx:205-223

Frozen spec v0.  For a state s (bitset over A undo actions):
    data_loss(s) = sum_{a not in s} p_a*size_a  +  sum_{a in s} (1-p_a)*size_a      [MB]
    downtime(s)  = sum_{a in s} cost_a                                               [s]
    score(s)     = -(data_loss + 0.1*downtime)
Spec v1 (round 2) adds the one NON-SEPARABLE term the reference's own candidate list implies
(threat-model.mdx:208-222: "Reverse file encryption" next to "Kill process python3"): reverting a file while the
process that encrypts it is still alive does not stick.  Every action a may name a GUARD g = guard_a in 0..31 -- the
index of the "kill process" action of the process that wrote it (guards live in state word 0) -- or -1:
    if a in s and guard_a >= 0 and guard_a not in s:   data_loss += w_a,   w_a = fl(p_g * u_a)
(with probability p_g, the process is malicious and re-encrypts the reverted file).  In the summation the term of an
applied action is vw_a = fl(v_a + w_a) instead of v_a -- one term per action as before, same order.  guard = None
(all -1) is spec v0 bit for bit.
All arithmetic is IEEE fp32 with a FIXED association order so that the CUDA
kernel can be bit-exact:
    NW     = 1, 2 or 4   (smallest with 1024*NW >= A);  chunk = 32*NW;  A_pad = 32*chunk
    lane l (0..31) sums its `chunk` consecutive actions a = l*chunk .. l*chunk+chunk-1
        sequentially in ascending a, starting from +0.0f  (padded actions add +0.0f)
    the 32 lane partials are combined by the xor-butterfly 1,2,4,8,16
        (== adjacent-pairs binary tree)
    score = -(dl + (0.1f * dt))        each op individually rounded (no FMA)
"""
import math
import numpy as np

F32 = np.float32


def layout(A):
    """-> (NW, chunk, A_pad, n_words)."""
    if A < 1 or A > 4096:
        raise ValueError("number of actions must be in 1..4096")
    NW = 1 if A <= 1024 else (2 if A <= 2048 else 4)
    return NW, 32 * NW, 1024 * NW, 32 * NW


def pad_actions(p, size, cost):
    p = np.asarray(p, dtype=F32); size = np.asarray(size, dtype=F32); cost = np.asarray(cost, dtype=F32)
    A = p.shape[0]
    _, _, A_pad, _ = layout(A)
    pp = np.zeros(A_pad, F32); ss = np.zeros(A_pad, F32); cc = np.zeros(A_pad, F32)
    pp[:A] = p; ss[:A] = size; cc[:A] = cost
    return pp, ss, cc


def action_terms(p, size, cost):
    """u_a = p*size (loss if NOT undone), v_a = (1-p)*size (loss if undone), fp32, padded."""
    pp, ss, cc = pad_actions(p, size, cost)
    u = (pp * ss).astype(F32)
    v = ((F32(1.0) - pp).astype(F32) * ss).astype(F32)
    return u, v, cc


def empty_state(A):
    """Bitset with the padding bits (a >= A) set, so they are never legal."""
    _, _, A_pad, nw = layout(A)
    bits = np.zeros(A_pad, dtype=bool)
    bits[A:] = True
    return pack_bits(bits[None, :])[0]


def pack_bits(bits):
    """bool [B, A_pad] -> uint32 [B, A_pad/32]; action a is bit (a & 31) of word (a >> 5)."""
    B, A_pad = bits.shape
    b = bits.reshape(B, A_pad // 32, 32).astype(np.uint32)
    sh = np.arange(32, dtype=np.uint32)
    return (b << sh).sum(axis=2, dtype=np.uint64).astype(np.uint32)


def unpack_bits(states):
    """uint32 [B, n_words] -> bool [B, 32*n_words]."""
    states = np.asarray(states, dtype=np.uint32)
    sh = np.arange(32, dtype=np.uint32)
    return (((states[:, :, None] >> sh) & np.uint32(1)) != 0).reshape(states.shape[0], -1)


def _lane_tree_sum(terms, chunk):
    """terms fp32 [B, A_pad] -> fp32 [B] in the spec'd order."""
    B = terms.shape[0]
    t = terms.reshape(B, 32, chunk)
    acc = np.zeros((B, 32), F32)
    for i in range(chunk):
        acc = (acc + t[:, :, i]).astype(F32)
    idx = np.arange(32)
    for s in (1, 2, 4, 8, 16):
        acc = (acc + acc[:, idx ^ s]).astype(F32)
    return acc[:, 0]


def guard_terms(p, size, cost, guard):
    """-> (g int64 [A_pad] (-1 = none), vw fp32 [A_pad] = fl(v_a + fl(p_g * u_a)) where a guard exists, else v_a)."""
    u, v, _ = action_terms(p, size, cost)
    A = np.asarray(p).shape[0]
    g = np.full(u.shape[0], -1, np.int64)
    if guard is not None:
        gg = np.asarray(guard, np.int64)
        if gg.shape[0] != A or (gg >= 32).any() or (gg >= A).any() or (gg < -1).any():
            raise ValueError("guard must be -1 or the index (< 32, < A) of the kill action")
        g[:A] = gg
    pg = np.where(g >= 0, np.asarray(p, F32)[np.maximum(g, 0).clip(max=A - 1)], F32(0.0)).astype(F32)
    w = (pg * u).astype(F32)
    vw = np.where(g >= 0, (v + w).astype(F32), v).astype(F32)
    return g, vw


def score(states, p, size, cost, guard=None):
    """states uint32 [B, n_words] (padding bits may be 0 or 1: padded terms are 0) -> fp32 [B]."""
    A = np.asarray(p).shape[0]
    NW, chunk, A_pad, nw = layout(A)
    states = np.asarray(states, dtype=np.uint32).reshape(-1, nw)
    u, v, c = action_terms(p, size, cost)
    applied = unpack_bits(states)
    if guard is not None:
        g, vw = guard_terms(p, size, cost, guard)
        guard_alive = (g[None, :] >= 0) & ~applied[:, np.maximum(g, 0)]          # the guard (kill) action is NOT in s
        v = np.where(guard_alive, vw[None, :], v[None, :]).astype(F32)
    else:
        v = v[None, :]
    dl_terms = np.where(applied, v, u[None, :]).astype(F32)
    dt_terms = np.where(applied, c[None, :], F32(0.0)).astype(F32)
    dl = _lane_tree_sum(dl_terms, chunk)
    dt = _lane_tree_sum(dt_terms, chunk)
    tenth = (F32(0.1) * dt).astype(F32)
    return (-((dl + tenth).astype(F32))).astype(F32)


def reward_bounds(p, size, cost, root_state=None, guard=None):
    """(lo, inv_range) fp32 normalisation constants used by the planner.

    lo / hi = worst / best score over all supersets of root_state, computed in
    float64 with exactly-rounded summation (math.fsum, order independent), then
    rounded to fp32.  value = (score - lo) * inv_range lies in ~[0, 1].
    """
    A = np.asarray(p).shape[0]
    u, v, c = action_terms(p, size, cost)
    u = u.astype(np.float64)[:A]; v = v.astype(np.float64)[:A]; c = c.astype(np.float64)[:A]
    applied_cost = v + 0.1 * c
    applied_worst = applied_cost
    if guard is not None:                      # an applied action may also pay its guard penalty: a valid (not tight) bound
        _, vw = guard_terms(p, size, cost, guard)
        applied_worst = vw.astype(np.float64)[:A] + 0.1 * c
    if root_state is None:
        fixed = np.zeros(A, bool)
    else:
        fixed = unpack_bits(np.asarray(root_state, np.uint32)[None, :])[0][:A]
    worst = np.where(fixed, applied_worst, np.maximum(u, applied_worst))
    best = np.where(fixed, applied_cost, np.minimum(u, applied_cost))
    lo = -math.fsum(worst.tolist())
    hi = -math.fsum(best.tolist())
    lo32 = F32(lo)
    rng = hi - float(lo32)
    inv = F32(1.0) if not (rng > 0.0) else F32(1.0 / rng)
    return lo32, inv
