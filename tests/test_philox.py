"""Philox4x32-10 known-answer tests (Random123 kat_vectors) for the numpy and C implementations, and the keyed
uniform the three implementations (numpy / C / CUDA) must share.  The CUDA side is checked in test_gpu_parity.py."""
import numpy as np

from oracle import philox_np

# (counter, key) -> output, from Random123's kat_vectors for philox4x32 with 10 rounds
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_numpy_kat():
    for ctr, key, exp in KAT:
        out = philox_np.philox4x32_10(*ctr, *key)
        assert tuple(int(x) for x in out) == exp


def test_c_kat(oracle_lib):
    for ctr, key, exp in KAT:
        out = oracle_lib.philox4x32_10(ctr, key)
        assert tuple(int(x) for x in out) == exp


def test_keyed_uniform_numpy_equals_c(oracle_lib):
    rs = np.random.RandomState(0)
    a = rs.randint(0, 2 ** 27, 200)
    b = rs.randint(0, 2 ** 27, 200)
    for seed in (0, 42, 2 ** 40 + 17):
        for day in (0, 5, 364):
            for stream in (0, 1, 2, 3):
                u = philox_np.keyed_uniform(seed, day, stream, a, b)
                assert ((u >= 0) & (u < 1)).all()
                for k in range(0, 200, 37):
                    assert u[k] == oracle_lib.keyed_uniform(seed, day, stream, int(a[k]), int(b[k]))


def test_keyed_uniform_is_53_bit_grid():
    u = philox_np.keyed_uniform(7, 3, 2, np.arange(1000), 11)
    assert np.array_equal(u * 2.0 ** 53, np.floor(u * 2.0 ** 53))
    assert abs(u.mean() - 0.5) < 0.05
