// Counter-based randomness for the epidemic step: Philox4x32-10 keyed on WHAT the draw is for.
//
// The reference draws np.random.random(N), random(N), random((U, I)) from the global MT19937 stream every day
// (contagious3.py:212, :224, :239).  Here u(day, stream, a, b) = 53-bit double from Philox4x32-10 with
// counter (a, b, day, stream) and key (seed_lo, seed_hi); a, b are ORIGINAL agent row indices.  The same
// function exists in numpy (oracle/philox_np.py) and C (oracle/dys_oracle.c) and is KAT-ed in tests/.
#pragma once
#include <stdint.h>

namespace dys {

enum : uint32_t { STREAM_ADMISSION = 0, STREAM_DEATH = 1, STREAM_INFECTION = 2, STREAM_COMPLIANCE = 3, STREAM_LOCKDOWN = 4 };

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                        uint32_t k0, uint32_t k1, uint32_t& o0, uint32_t& o1,
                                                        uint32_t& o2, uint32_t& o3)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o0 = c0; o1 = c1; o2 = c2; o3 = c3;
}

__host__ __device__ __forceinline__ double keyed_uniform(uint64_t seed, uint32_t day, uint32_t stream, uint32_t a,
                                                          uint32_t b)
{
    uint32_t o0, o1, o2, o3;
    philox4x32_10(a, b, day, stream, (uint32_t)seed, (uint32_t)(seed >> 32), o0, o1, o2, o3);
    const uint64_t w = (uint64_t)o0 | ((uint64_t)o1 << 32);
    return (double)(w >> 11) * (1.0 / 9007199254740992.0);   // exact: 53-bit integer times 2^-53
}

}  // namespace dys
