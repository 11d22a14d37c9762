"""Lock-step emulation of warp-level integer code in NumPy: every "register" is a uint32 array over the
32 lanes.  Used to check the lane layouts / shuffles / byte permutes of the v2 kernels against the oracle
on the CPU before spending GPU time (the CUDA sources are transliterated line by line)."""
import numpy as np

LANES = np.arange(32)
U = np.uint32


def u32(x):
    return (np.asarray(x).astype(np.int64) & 0xFFFFFFFF).astype(np.uint32)


def shfl_xor(x, m):
    return x[LANES ^ m]


def shfl(x, idx):
    return x[np.asarray(idx) & 31]


def byte_perm(a, b, sel):
    a, b = u32(a).astype(np.uint64), u32(b).astype(np.uint64)
    if np.ndim(a) == 0:
        a = np.full(32, a, np.uint64)
    if np.ndim(b) == 0:
        b = np.full(32, b, np.uint64)
    src = a | (b << np.uint64(32))
    out = np.zeros(32, np.uint64)
    for i in range(4):
        n = (sel >> (4 * i)) & 0xF
        assert n < 8, "sign-replicating selectors are not emulated"
        out |= ((src >> np.uint64(8 * n)) & np.uint64(0xFF)) << np.uint64(8 * i)
    return out.astype(np.uint32)


def mul(a, k):
    return u32(a.astype(np.uint64) * np.asarray(k).astype(np.uint64))


def add(*xs):
    s = np.zeros(32, np.uint64)
    for x in xs:
        s = s + np.asarray(x).astype(np.uint64)
    return u32(s)
