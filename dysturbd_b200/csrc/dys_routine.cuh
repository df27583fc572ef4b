// Routine generation on the device (contagious3.py:139-170): every agent's static visit list -- home,
// then before each anchor (agents[:,12], agents[:,18], home) up to k flexible stops, each taken on a coin flip (:150) and
// drawn uniformly (:155) from
//     (buildings within bld_dist TO the anchor  ∩  buildings within bld_dist FROM the current position)  ∪
//     (buildings within a_dist of the agent),
// enumerated in ascending building-id order (np.intersect1d / np.union1d sort their result).  The three sets arrive as
// sorted index lists over the buildings renumbered by ascending id (dysturbd_b200/routines.py derives them from the
// reference's shortest-path matrix; a sparse network builder provides them directly).  One thread per agent: the agent's
// stops depend on each other, the agents do not -- the draws are keyed on (agent row, coin index), not on a global stream.
#pragma once
#include "dys_common.cuh"

namespace dys {

enum : uint32_t { STREAM_ROUTINE_COIN = 5, STREAM_ROUTINE_CHOICE = 6 };

// |(A ∩ B) ∪ C| of three ascending lists, or (pick >= 0) its pick-th element in ascending order
__device__ __forceinline__ int64_t union_pick(const int32_t* __restrict__ A, const int64_t na, const int32_t* __restrict__ B,
                                              const int64_t nb, const int32_t* __restrict__ C, const int64_t nc, const int64_t pick,
                                              int32_t& out)
{
    constexpr int64_t INF = (int64_t)1 << 40;
    int64_t pa = 0, pb = 0, pc = 0, count = 0;
    for (;;) {
        while (pa < na && pb < nb) {
            const int32_t x = __ldg(A + pa), y = __ldg(B + pb);
            if (x == y) break;
            if (x < y) ++pa; else ++pb;
        }
        const int64_t x = (pa < na && pb < nb) ? (int64_t)__ldg(A + pa) : INF;
        const int64_t y = pc < nc ? (int64_t)__ldg(C + pc) : INF;
        const int64_t v = x < y ? x : y;
        if (v == INF) break;
        if (count == pick) { out = (int32_t)v; return count; }
        ++count;
        if (x == v) { ++pa; ++pb; }
        if (y == v) ++pc;
    }
    return count;
}

__global__ void __launch_bounds__(128) k_routines(const int64_t N, const int V, const int k, const int32_t* __restrict__ home,
                                                  const int32_t* __restrict__ anchors, const int64_t* __restrict__ out_ptr,
                                                  const int32_t* __restrict__ out_idx, const int64_t* __restrict__ in_ptr,
                                                  const int32_t* __restrict__ in_idx, const int64_t* __restrict__ near_ptr,
                                                  const int32_t* __restrict__ near_idx, const uint64_t seed,
                                                  int32_t* __restrict__ routine)
{
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= N) return;
    int32_t* r = routine + a * V;
    for (int s = 0; s < V; ++s) r[s] = -1;
    int32_t cur = home[a];
    int nv = 0, q = 0;
    r[nv++] = cur;
    const int32_t* C = near_idx + near_ptr[a];
    const int64_t nc = near_ptr[a + 1] - near_ptr[a];
    for (int s = 0; s < 3; ++s) {
        const int32_t i = anchors[a * 3 + s];
        if (i < 0) continue;                                                              // NaN anchor, :147
        for (int j = 0; j < k; ++j, ++q) {
            if (!(keyed_uniform(seed, 0u, STREAM_ROUTINE_COIN, (uint32_t)a, (uint32_t)q) >= 0.5)) continue;      // :150
            const int32_t* A = in_idx + in_ptr[i];
            const int32_t* B = out_idx + out_ptr[cur];
            const int64_t na = in_ptr[i + 1] - in_ptr[i], nb = out_ptr[cur + 1] - out_ptr[cur];
            int32_t dest = cur;
            const int64_t len = union_pick(A, na, B, nb, C, nc, -1, dest);
            const double u = keyed_uniform(seed, 0u, STREAM_ROUTINE_CHOICE, (uint32_t)a, (uint32_t)q);
            union_pick(A, na, B, nb, C, nc, (int64_t)__dmul_rn(u, (double)len), dest);                           // :155
            cur = dest;
            r[nv++] = cur;
        }
        cur = i;                                                                          // :160-161
        r[nv++] = cur;
    }
}

// every list ascending and inside [0, n_bld); anchors / home inside the buildings
__global__ void k_validate_lists(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx, const int64_t rows,
                                 const int64_t n_bld, int32_t* bad)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    if (r == 0 && ptr[0] != 0) atomicOr(bad, 1);
    if (ptr[r + 1] < ptr[r]) { atomicOr(bad, 1); return; }
    for (int64_t e = ptr[r]; e < ptr[r + 1]; ++e) {
        if (idx[e] < 0 || idx[e] >= n_bld) atomicOr(bad, 2);
        if (e > ptr[r] && idx[e] <= idx[e - 1]) atomicOr(bad, 4);
    }
}

}  // namespace dys
