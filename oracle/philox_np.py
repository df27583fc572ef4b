"""TEST INFRASTRUCTURE (oracle) -- numpy Philox4x32-10 and the keyed uniform u(day, stream, a, b).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

The reference draws from the global, unseeded numpy MT19937 in a fixed order each day
(/root/reference/contagious3.py:212 `random(N)`, :224 `random(N)`, :239 `random((U,I))`).
For parity we replace those three arrays by a counter-based generator keyed on *what the draw
is for* instead of *when it was drawn*:

    u(day, stream, a, b) = ((o0 | o1 << 32) >> 11) * 2**-53
    (o0, o1, o2, o3)     = Philox4x32-10(counter = (a, b, day, stream), key = (seed_lo, seed_hi))

stream 0 = admission draw of agent a (b = 0), stream 1 = death draw of agent a (b = 0),
stream 2 = infection draw for the pair (susceptible a, infectious b) in ORIGINAL row indices,
stream 3 = quarantine-compliance draw of agent a (extension; unused at compliance 1.0).
The 53-bit construction is the one numpy's Generator.random uses.  The identical function is
implemented in C (oracle/dys_oracle.c) and CUDA (dysturbd_b200/csrc/philox.cuh); the three are
KAT-ed against the Random123 known-answer vectors and against each other in tests/.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)
S32 = np.uint64(32)

STREAM_ADMISSION = 0
STREAM_DEATH = 1
STREAM_INFECTION = 2
STREAM_COMPLIANCE = 3


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  Counter words are broadcastable integer arrays, key words python ints.
    Returns four uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(*(np.asarray(c, dtype=np.uint64) & MASK32 for c in (c0, c1, c2, c3)))
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for r in range(10):
        p0 = M0 * c0          # 32x32 -> 64 bit products, exact in uint64
        p1 = M1 * c2
        hi0, lo0 = p0 >> S32, p0 & MASK32
        hi1, lo1 = p1 >> S32, p1 & MASK32
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def keyed_uniform(seed, day, stream, a, b=0):
    """u(day, stream, a, b) in [0, 1) as float64; a, b broadcastable integer arrays."""
    seed = int(seed)
    o0, o1, _, _ = philox4x32_10(a, b, day, stream, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    w = o0.astype(np.uint64) | (o1.astype(np.uint64) << S32)
    return (w >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)
