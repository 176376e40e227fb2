"""Pin the oracle's Philox4x32-10 against the Random123 known-answer vectors."""
import numpy as np

from oracle.philox import philox4x32_10

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_random123_kat():
    for ctr, key, want in KAT:
        got = philox4x32_10(*ctr, *key)
        assert tuple(int(g) for g in got) == want


def test_vectorised_matches_scalar():
    r = np.arange(17, dtype=np.uint32)
    out = philox4x32_10(r, 3, 5, 0, 11, 12)
    for i in range(17):
        s = philox4x32_10(int(r[i]), 3, 5, 0, 11, 12)
        assert all(int(out[j][i]) == int(s[j]) for j in range(4))
