// Network construction on the device: the two O(N^2) steps that precede the day loop.
//
//  * k_sim_edges   -- the agent-agent edges of communities.create_network (communities.py:22-40, 65-70): every ordered pair is
//                     scored as the reference scores it (same float64 operations in the same order, no fused multiply-add),
//                     pairs above the threshold come out as CSR rows in np.where order.  No N x N matrix exists anywhere:
//                     a block keeps 8 rows and walks the columns in shared-memory tiles.
//  * k_bld_edges   -- the building-building edges of the same function (the gravity model, communities.py:12-21, 53-54), with
//                     numpy's pairwise row sums reproduced addition by addition (k_bld_rowsum).
//  * k_sp_relax    -- contagious3.py:124-130, scipy.sparse.csgraph.shortest_path: label-correcting relaxation for a batch of
//                     sources at once, D[vertex][source] with the sources contiguous (a warp relaxes one vertex for 128 sources,
//                     four per lane: every neighbour read is 1 KB contiguous).  Distances are the left-to-right sums d(v) = fl(d(u) + w(u,v));
//                     the least fixed point of that recurrence is unique (rounded addition is monotone and fl(d + w) >= d for
//                     w >= 0), so the result is bit-identical to Dijkstra's whatever order the relaxations happen in.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace dys {

constexpr int NET_BLOCK = 256;
constexpr int NET_ROWS = NET_BLOCK / 32;      // rows per block of k_sim_edges (one warp each)
constexpr int NET_TILE = 256;                 // columns staged per step

__device__ __forceinline__ void atomic_max_double_nonneg(unsigned long long* addr, const double v)
{
    atomicMax(addr, (unsigned long long)__double_as_longlong(v));      // (non-negative doubles order like their bit patterns)
}

// max over pairs of dx^2 + dy^2 (the largest distance is its sqrt: sqrt is monotone) -- communities.py:36 `np.max(agent_dists)`
__global__ void __launch_bounds__(NET_BLOCK) k_max_d2(const double* __restrict__ x, const double* __restrict__ y, const int64_t n,
                                                      unsigned long long* __restrict__ out_bits)
{
    __shared__ double sx[NET_TILE], sy[NET_TILE];
    __shared__ double s_red[NET_BLOCK / 32];
    const int64_t i = (int64_t)blockIdx.x * NET_BLOCK + threadIdx.x;
    const double xi = i < n ? x[i] : 0.0, yi = i < n ? y[i] : 0.0;
    double best = 0.0;
    for (int64_t j0 = 0; j0 < n; j0 += NET_TILE) {
        const int64_t j = j0 + threadIdx.x;
        sx[threadIdx.x] = j < n ? x[j] : xi;
        sy[threadIdx.x] = j < n ? y[j] : yi;
        __syncthreads();
        if (i < n) {
            const int m = (int)((n - j0) < NET_TILE ? (n - j0) : NET_TILE);
            for (int k = 0; k < m; ++k) {
                const double dx = __dsub_rn(sx[k], xi), dy = __dsub_rn(sy[k], yi);
                const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                best = d2 > best ? d2 : best;
            }
        }
        __syncthreads();
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
    }
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < NET_BLOCK / 32; ++k) best = s_red[k] > best ? s_red[k] : best;
        atomic_max_double_nonneg(out_bits, best);
    }
}

struct SimParams {
    int64_t n;
    double max_age, max_inc, max_dist;       // the three normalisers (communities.py:27, 30, 36)
    double w_a, w_i, w_d, min_prob;
    int32_t prefilter;                       // 1: pre_* bound the pairs that can pass (weights in (0, 1], finite positive normalisers)
    double pre_a, pre_i, pre_d2;
};

// FILL = false: count[i] = edges of row i.  FILL = true: col / prob at rowptr[i] + rank, columns ascending.
template <bool FILL>
__global__ void __launch_bounds__(NET_BLOCK) k_sim_edges(const SimParams q, const double* __restrict__ age, const double* __restrict__ inc,
                                                         const double* __restrict__ x, const double* __restrict__ y,
                                                         const int64_t* __restrict__ hh, int64_t* __restrict__ count,
                                                         const int64_t* __restrict__ rowptr, int32_t* __restrict__ col,
                                                         double* __restrict__ prob)
{
    __shared__ double s_age[NET_TILE], s_inc[NET_TILE], s_x[NET_TILE], s_y[NET_TILE];
    __shared__ int64_t s_hh[NET_TILE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * NET_ROWS + warp;
    const bool row = i < q.n;
    const double age_i = row ? age[i] : 0.0, inc_i = row ? inc[i] : 0.0, x_i = row ? x[i] : 0.0, y_i = row ? y[i] : 0.0;
    const int64_t hh_i = row ? hh[i] : -1;
    int64_t k = FILL && row ? rowptr[i] : 0;
    for (int64_t j0 = 0; j0 < q.n; j0 += NET_TILE) {
        const int64_t jt = j0 + threadIdx.x;
        if (jt < q.n) {
            s_age[threadIdx.x] = age[jt]; s_inc[threadIdx.x] = inc[jt]; s_x[threadIdx.x] = x[jt]; s_y[threadIdx.x] = y[jt];
            s_hh[threadIdx.x] = hh[jt];
        }
        __syncthreads();
        if (row) {
            for (int t = 0; t < NET_TILE; t += 32) {
                const int64_t j = j0 + t + lane;
                bool edge = false;
                double p = 0.0;
                if (j < q.n && j != i) {
                    const int s = t + lane;
                    if (s_hh[s] == hh_i) {
                        p = 1.0;                                                                // communities.py:39
                        edge = p > q.min_prob;
                    } else {
                        const double da = fabs(__dsub_rn(s_age[s], age_i));                     // :29
                        const double di = fabs(__dsub_rn(s_inc[s], inc_i));                     // :24
                        const double dx = __dsub_rn(s_x[s], x_i), dy = __dsub_rn(s_y[s], y_i);
                        const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));      // :35 (minkowski, p = 2)
                        // every factor is <= 1 and so are the weights: a pair whose age, income or distance factor alone is
                        // below the threshold cannot pass (rounded products are monotone).  The bounds carry a 0.1 % margin.
                        if (!(q.prefilter && (da > q.pre_a || di > q.pre_i || d2 > q.pre_d2))) {
                            const double fa = __dsub_rn(1.0, __ddiv_rn(da, q.max_age));         // :30
                            const double fi = __dsub_rn(1.0, __ddiv_rn(di, q.max_inc));         // :27
                            const double fd = __dsub_rn(1.0, __ddiv_rn(__dsqrt_rn(d2), q.max_dist));  // :36
                            p = __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(q.w_a, fa), q.w_i), fi), q.w_d), fd);   // :38
                            edge = p > q.min_prob;                                              // :65
                        }
                    }
                }
                const unsigned m = __ballot_sync(0xffffffffu, edge);
                if (FILL && edge) {
                    const int64_t at = k + __popc(m & ((1u << lane) - 1u));
                    col[at] = (int32_t)j;
                    prob[at] = p;
                }
                k += __popc(m);
            }
        }
        __syncthreads();
    }
    if (!FILL && row && lane == 0) count[i] = k;
}

// ---- building-building edges (communities.py:12-21, 53-54): the gravity model ----
// scores[i][j] = fs[j] / dist(i,j)^2 with dist^2 the SQUARE OF THE ROUNDED DISTANCE (`distance_matrix(...)**beta`, beta = 2), a zero
// replaced by 0.00001 (:14) and scores[i][i] = 0 (:17); prob = scores / row sum (:19); an edge wherever prob > min_prob, i != j.
__device__ __forceinline__ double bld_score(const double fs_j, const double x_i, const double y_i, const double x_j, const double y_j,
                                            const bool self)
{
    const double dx = __dsub_rn(x_j, x_i), dy = __dsub_rn(y_j, y_i);
    const double d = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    double d2 = __dmul_rn(d, d);
    if (d2 == 0.0) d2 = 0.00001;
    return self ? 0.0 : __ddiv_rn(fs_j, d2);
}

// The row sum is numpy's `scores.sum(axis=1)`: PAIRWISE summation of the contiguous row.  Its shape depends on B alone, so the
// host flattens numpy's recursion into a program -- LEAF(j0, n <= 128): push the sum of that block (eight strided accumulators,
// combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the remainder; fewer than 8 elements: a plain loop); ADD: pop two, push
// their sum -- and every row (one thread each) runs the same program: no divergence, operand loads are warp-uniform.
struct SumOp { int32_t j0, n; };              // n > 0: LEAF;  n == 0: ADD
__global__ void __launch_bounds__(NET_BLOCK) k_bld_rowsum(const int64_t B, const double* __restrict__ fs, const double* __restrict__ x,
                                                          const double* __restrict__ y, const SumOp* __restrict__ prog, const int n_ops,
                                                          double* __restrict__ row_sum)
{
    const int64_t i = (int64_t)blockIdx.x * NET_BLOCK + threadIdx.x;
    if (i >= B) return;
    const double x_i = x[i], y_i = y[i];
    double stack[40];
    int sp = 0;
    for (int op = 0; op < n_ops; ++op) {
        const SumOp o = prog[op];
        if (o.n == 0) {
            --sp;
            stack[sp - 1] = __dadd_rn(stack[sp - 1], stack[sp]);
            continue;
        }
        auto score = [&](const int64_t j) { return bld_score(fs[j], x_i, y_i, x[j], y[j], j == i); };
        double res;
        if (o.n < 8) {
            res = 0.0;
            for (int k = 0; k < o.n; ++k) res = __dadd_rn(res, score(o.j0 + k));
        } else {
            double r[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) r[q] = score(o.j0 + q);
            int k = 8;
            for (; k < o.n - (o.n % 8); k += 8) {
#pragma unroll
                for (int q = 0; q < 8; ++q) r[q] = __dadd_rn(r[q], score(o.j0 + k + q));
            }
            res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])), __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
            for (; k < o.n; ++k) res = __dadd_rn(res, score(o.j0 + k));
        }
        stack[sp++] = res;
    }
    row_sum[i] = stack[0];
}

template <bool FILL>
__global__ void __launch_bounds__(NET_BLOCK) k_bld_edges(const int64_t B, const double min_prob, const double* __restrict__ fs,
                                                         const double* __restrict__ x, const double* __restrict__ y,
                                                         const double* __restrict__ row_sum, int64_t* __restrict__ count,
                                                         const int64_t* __restrict__ rowptr, int32_t* __restrict__ col,
                                                         double* __restrict__ prob)
{
    __shared__ double s_fs[NET_TILE], s_x[NET_TILE], s_y[NET_TILE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * NET_ROWS + warp;
    const bool row = i < B;
    const double x_i = row ? x[i] : 0.0, y_i = row ? y[i] : 0.0, sum_i = row ? row_sum[i] : 1.0;
    int64_t k = FILL && row ? rowptr[i] : 0;
    for (int64_t j0 = 0; j0 < B; j0 += NET_TILE) {
        const int64_t jt = j0 + threadIdx.x;
        if (jt < B) { s_fs[threadIdx.x] = fs[jt]; s_x[threadIdx.x] = x[jt]; s_y[threadIdx.x] = y[jt]; }
        __syncthreads();
        if (row) {
            for (int t = 0; t < NET_TILE; t += 32) {
                const int64_t j = j0 + t + lane;
                bool edge = false;
                double p = 0.0;
                if (j < B && j != i) {
                    p = __ddiv_rn(bld_score(s_fs[t + lane], x_i, y_i, s_x[t + lane], s_y[t + lane], false), sum_i);     // :19
                    edge = p > min_prob;                                                                               // :53
                }
                const unsigned m = __ballot_sync(0xffffffffu, edge);
                if (FILL && edge) {
                    const int64_t at = k + __popc(m & ((1u << lane) - 1u));
                    col[at] = (int32_t)j;
                    prob[at] = p;
                }
                k += __popc(m);
            }
        }
        __syncthreads();
    }
    if (!FILL && row && lane == 0) count[i] = k;
}

// ---- shortest paths ----
constexpr int SP_CHUNK = 128;                 // sources per warp of k_sp_relax (four per lane)
// D[v * batch + s]: distance from source s of the batch to vertex v
__global__ void __launch_bounds__(NET_BLOCK) k_sp_init(double* __restrict__ D, const int64_t n, const int batch,
                                                       const int32_t* __restrict__ sources, const int n_src, uint8_t* __restrict__ dirty,
                                                       const int marks_per_vertex)
{
    const int64_t at = (int64_t)blockIdx.x * NET_BLOCK + threadIdx.x;
    if (at >= n * batch) return;
    const int64_t v = at / batch;
    const int s = (int)(at - v * batch);
    const bool is_src = s < n_src && sources[s] == v;
    D[at] = is_src ? 0.0 : __longlong_as_double(0x7ff0000000000000ll);
    if (is_src) dirty[v * marks_per_vertex + (marks_per_vertex > 1 ? s / SP_CHUNK : 0)] = 1;
}

// One warp = vertex v x 128 sources, four per lane (two 16-byte loads per neighbour and lane: the per-edge work -- index,
// mark, weight -- is shared by four relaxations; the ncu capture of the one-source-per-lane version showed the kernel bound by
// instruction issue at 6 % of the L2 throughput).  A vertex is looked at only if an in-neighbour changed in the previous
// sweep, and only those neighbours are read.  The "changed" marks are kept per (vertex, 128 sources) -- marks_per_vertex =
// batch / 128 -- so a warp only follows the wave fronts of its own sources (marks_per_vertex = 1: one mark per vertex for the
// whole batch).  Updates are in place: a value read here is a real path length at all times, so reading one that another
// warp is just improving is harmless (the neighbour is marked again and v comes back in the next sweep).
__global__ void __launch_bounds__(NET_BLOCK) k_sp_relax(const int64_t n, const int batch, const int64_t* __restrict__ in_ptr,
                                                        const int32_t* __restrict__ in_src, const double* __restrict__ in_w,
                                                        double* D, const uint8_t* __restrict__ dirty_prev,
                                                        uint8_t* __restrict__ dirty_next, int32_t* __restrict__ changed,
                                                        const int marks_per_vertex, const double bound)
{
    // bound: only distances BELOW it are wanted (+inf: all).  A label that is not below the bound is never stored, so the wave
    // fronts stop at that radius; exact for every entry below the bound, because weights are >= 0 and rounded sums monotone:
    // a path through a vertex at or beyond the bound cannot come back under it.
    const int chunks = batch / SP_CHUNK;
    const int64_t wid = ((int64_t)blockIdx.x * NET_BLOCK + threadIdx.x) >> 5;
    if (wid >= n * chunks) return;
    const int lane = threadIdx.x & 31;
    const int64_t v = wid / chunks;
    const int c = (int)(wid - v * chunks);
    const int64_t e0 = in_ptr[v], e1 = in_ptr[v + 1];
    const int64_t fs = marks_per_vertex;
    const int fo = marks_per_vertex > 1 ? c : 0;
    bool any = false;
    for (int64_t e = e0 + lane; e < e1; e += 32) any |= dirty_prev[in_src[e] * fs + fo] != 0;
    if (!__any_sync(0xffffffffu, any)) return;
    const int64_t off = (int64_t)c * SP_CHUNK + lane * 4;
    double2* mine = reinterpret_cast<double2*>(D + v * batch + off);
    const double2 o0 = __ldcg(mine), o1 = __ldcg(mine + 1);
    double2 b0 = o0, b1 = o1;
    for (int64_t e = e0; e < e1; ++e) {
        const int32_t u = in_src[e];
        if (!dirty_prev[u * fs + fo]) continue;
        const double w = in_w[e];
        const double2* theirs = reinterpret_cast<const double2*>(D + (int64_t)u * batch + off);
        const double2 x0 = __ldcg(theirs), x1 = __ldcg(theirs + 1);
        const double c0 = __dadd_rn(x0.x, w), c1 = __dadd_rn(x0.y, w), c2 = __dadd_rn(x1.x, w), c3 = __dadd_rn(x1.y, w);
        b0.x = (c0 < b0.x && c0 < bound) ? c0 : b0.x;
        b0.y = (c1 < b0.y && c1 < bound) ? c1 : b0.y;
        b1.x = (c2 < b1.x && c2 < bound) ? c2 : b1.x;
        b1.y = (c3 < b1.y && c3 < bound) ? c3 : b1.y;
    }
    const bool better = b0.x < o0.x || b0.y < o0.y || b1.x < o1.x || b1.y < o1.y;
    if (better) { __stcg(mine, b0); __stcg(mine + 1, b1); }
    if (__any_sync(0xffffffffu, better) && lane == 0) {
        dirty_next[v * fs + fo] = 1;
        *changed = 1;
    }
}

// out[s][v] = D[v][s] (row-major rows per source, what shortest_path returns)
__global__ void __launch_bounds__(NET_BLOCK) k_sp_transpose(const double* __restrict__ D, const int64_t n, const int batch, const int n_src,
                                                            double* __restrict__ out)
{
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;           // 32 x 8
    const int64_t v0 = (int64_t)blockIdx.x * 32;
    const int s0 = blockIdx.y * 32;
    for (int r = ty; r < 32; r += NET_BLOCK / 32) {
        const int64_t v = v0 + r;
        tile[r][tx] = (v < n && s0 + tx < batch) ? D[v * batch + s0 + tx] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += NET_BLOCK / 32) {
        const int s = s0 + r;
        const int64_t v = v0 + tx;
        if (s < n_src && v < n) out[(int64_t)s * n + v] = tile[tx][r];
    }
}

// rows[s][v] (what k_sp_transpose wrote) -> the vertices of every source that are nearer than `bound` (+inf: every reachable
// one) and, if col_keep is given, marked in it, as CSR: FILL = false counts, FILL = true
// writes (vertex, distance) at ptr[s] + rank, vertices ascending.  skip_self leaves the source itself out (the reference
// sets the diagonal to inf before it takes exp(-dists), contagious3.py:126-127).
template <bool FILL>
__global__ void __launch_bounds__(NET_BLOCK) k_sp_compact(const double* __restrict__ rows, const int64_t n, const int n_src,
                                                          const int32_t* __restrict__ sources, const int skip_self,
                                                          const double bound, const uint8_t* __restrict__ col_keep,
                                                          int64_t* __restrict__ count, const int64_t* __restrict__ ptr,
                                                          int32_t* __restrict__ col, double* __restrict__ dist)
{
    const int lane = threadIdx.x & 31;
    const int s = (int)(((int64_t)blockIdx.x * NET_BLOCK + threadIdx.x) >> 5);
    if (s >= n_src) return;
    const double* row = rows + (int64_t)s * n;
    const int64_t self = skip_self ? (int64_t)sources[s] : -1;
    int64_t k = FILL ? ptr[s] : 0;
    for (int64_t v0 = 0; v0 < n; v0 += 32) {
        const int64_t v = v0 + lane;
        const double d = v < n ? row[v] : __longlong_as_double(0x7ff0000000000000ll);
        const bool keep = d < bound && v != self && (col_keep == nullptr || col_keep[v < n ? v : 0] != 0);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (FILL && keep) {
            const int64_t at = k + __popc(m & ((1u << lane) - 1u));
            col[at] = (int32_t)v;
            dist[at] = d;
        }
        k += __popc(m);
    }
    if (!FILL && lane == 0) count[s] = k;
}

}  // namespace dys
