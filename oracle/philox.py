"""Philox4x32-10 counter-based RNG, vectorised numpy restatement (oracle; test infra).

Algorithm: Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11),
the Random123 ``philox4x32_R(10, ctr, key)`` function.  The reference names no
RNG (SURVEY.md 7.3 asks for a counter-based generator implemented identically
on both sides so MCTS action indices are bit-exact); this is that generator.
Pinned by the Random123 known-answer vectors in tests/test_oracle_philox.py.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """All arguments broadcastable uint32 arrays / ints.  Returns 4 uint32 arrays."""
    c0, c1, c2, c3, k0, k1 = np.broadcast_arrays(
        *[np.asarray(v, dtype=np.uint32) for v in (c0, c1, c2, c3, k0, k1)])
    c0 = c0.copy(); c1 = c1.copy(); c2 = c2.copy(); c3 = c3.copy()
    k0 = k0.copy(); k1 = k1.copy()
    with np.errstate(over="ignore"):
        for rnd in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0 = (p0 >> _S32).astype(np.uint32); lo0 = (p0 & _MASK).astype(np.uint32)
            hi1 = (p1 >> _S32).astype(np.uint32); lo1 = (p1 & _MASK).astype(np.uint32)
            n0 = hi1 ^ c1 ^ k0
            n1 = lo1
            n2 = hi0 ^ c3 ^ k1
            n3 = lo0
            c0, c1, c2, c3 = n0, n1, n2, n3
            if rnd != 9:
                k0 = (k0 + W0).astype(np.uint32)
                k1 = (k1 + W1).astype(np.uint32)
    return c0, c1, c2, c3
