"""Host-side twin of csrc/philox.cuh: u(day, stream, a, b) = 53-bit double from Philox4x32-10 with counter (a, b, day,
stream) and key (seed_lo, seed_hi).  Only the host lockdown policy in keyed mode uses it (one draw per building on the
rare day a GRADUAL scenario closes half of a class); the per-agent and per-pair draws are made on the device."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
STREAM_LOCKDOWN = 4


def keyed_uniform(seed, day, stream, a, b=0):
    a = np.atleast_1d(np.asarray(a, dtype=np.uint64))
    c0 = a & MASK
    c1 = np.broadcast_to(np.asarray(b, dtype=np.uint64) & MASK, a.shape).copy()
    c2 = np.full(a.shape, int(day) & 0xFFFFFFFF, dtype=np.uint64)
    c3 = np.full(a.shape, int(stream) & 0xFFFFFFFF, dtype=np.uint64)
    k0, k1 = int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, p1 & MASK, n2, p0 & MASK
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    w = c0 | (c1 << np.uint64(32))
    return (w >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
