// Load-time kernels: input validation, the int32 <-> int16 day columns, and the transposed, class-annotated
// interaction_prob that k_push walks.  None of this runs per day.
#pragma once
#include "dys_common.cuh"

namespace dys {

// int32 day column -> int16 field `field` (0 t_inf, 1 t_q, 2 t_adm) of the agent records, with range check
__global__ void k_narrow_days(const int32_t* __restrict__ in, AgentRec* __restrict__ rec, int field, int64_t n, int32_t* bad)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t v = in[i];
    if (v < -32768 || v > 32767) atomicOr(bad, 256);
    reinterpret_cast<int16_t*>(rec + i)[field] = (int16_t)v;
}

__global__ void k_widen_days(const AgentRec* __restrict__ rec, int field, int32_t* __restrict__ out, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = reinterpret_cast<const int16_t*>(rec + i)[field];
}

// every index inside its range; t_adm == 0 wherever status is in {1, 2, 6, 7} (contagious3.py:259 keeps that invariant);
// a susceptible agent has no infector yet (agents[:,21] == 0 until :248 writes it)
__global__ void k_validate_population(const DevCtx c, int32_t* bad)
{
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= c.n_loc) return;
    const uint8_t s = c.status[a];
    const bool known = s == ST_S || s == ST_I || s == ST_QH || s == ST_QI || s == ST_D || s == ST_H || s == ST_R || s == ST_X;
    if (!known) atomicOr(bad, 1);
    const AgentRec r = c.rec[a];
    if (c.hh[a] < 0 || c.hh[a] >= c.H || r.home < 0 || r.home >= c.B_out || r.sa < -1 || r.sa >= c.S_out) atomicOr(bad, 2);
    if ((s == ST_S || s == ST_I || s == ST_R || s == ST_X) && r.t_adm != 0) atomicOr(bad, 4);
    if (c.src[a] < -1 || c.src[a] >= c.n_glob) atomicOr(bad, 8);
    if ((s == ST_S || s == ST_QH) && c.src[a] != -1) atomicOr(bad, 512);
    // a recovered agent is only idle (skipped by the sweep) once it has left the 21-bin days-since-infection histogram;
    // agents that recover during the run always have (recover >= 21), agents LOADED as recovered may not
    if (s == ST_R) atomicMax(bad + 1, (int32_t)r.t_inf + HIST_LEN);
}

// dys_load_quirk: trigger lists must be monotone and name owned households (raise_household stores qflag[trig_hh[t]])
__global__ void k_validate_quirk(const int64_t* __restrict__ trig_ptr, const int32_t* __restrict__ trig_hh, int64_t n, int64_t nt,
                                 int64_t H, int32_t* bad)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && trig_ptr[0] != 0) atomicOr(bad, 2048);
    if (i < n && (trig_ptr[i + 1] < trig_ptr[i] || trig_ptr[i + 1] > nt)) atomicOr(bad, 2048);
    if (i < nt && (trig_hh[i] < 0 || trig_hh[i] >= H)) atomicOr(bad, 4096);
}

// household / home building / area indices arrive global and are stored relative to the owned ranges
__global__ void k_localize(int32_t* __restrict__ hh, const int32_t* __restrict__ home, const int32_t* __restrict__ sa,
                           AgentRec* __restrict__ rec, int64_t n, int32_t hh_lo, int32_t bld_lo, int32_t area_lo)
{
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    hh[a] -= hh_lo;
    rec[a].home = home[a] - bld_lo;
    const int32_t s = sa[a];
    rec[a].sa = s >= 0 ? (s >= area_lo ? s - area_lo : -2) : s;         // -2: outside the owned areas -> rejected by validation
}

// household -> members: counts, then (after k_scan_counts) the member lists; order inside a household is irrelevant
__global__ void k_count_members(const int32_t* __restrict__ hh, int64_t n, int32_t* __restrict__ cnt)
{
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a < n) atomicAdd(&cnt[hh[a]], 1);
}

__global__ void k_fill_members(const int32_t* __restrict__ hh, int64_t n, unsigned long long* __restrict__ cursor,
                               int32_t* __restrict__ mem)
{
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a < n) mem[atomicAdd(&cursor[hh[a]], 1ull)] = (int32_t)a;
}

// household, member list position and admission probability into the agent records
__global__ void k_pack_records(const int32_t* __restrict__ hh, const int64_t* __restrict__ hh_ptr, const double* __restrict__ p_adm,
                               AgentRec* __restrict__ rec, int64_t n, int32_t* bad)
{
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const int32_t h = hh[a];
    const int64_t first = hh_ptr[h], size = hh_ptr[h + 1] - first;
    if (size > 32767) atomicOr(bad, 1024);
    rec[a].hh = h;
    rec[a].hh_slot = (int32_t)first;
    rec[a].hh_size = (int16_t)size;
    rec[a].p_adm = p_adm[a];
}

__global__ void k_validate_routine(const int32_t* __restrict__ routine, int64_t n, int64_t B, int32_t* bad)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && (routine[i] < -1 || routine[i] >= B)) atomicOr(bad, 16);
}

// Routines are static (computed once per simulation, contagious3.py:139-170), so WHICH slots two agents share is static
// too.  Four bits per interaction_prob non-zero (u = row = susceptible side, i = column = infectious side):
//   HH  home(u) == home(i)                  HO  home(u) is a non-home stop of i
//   OH  home(i) is a non-home stop of u     OO  they share a non-home stop
// With every building open, exposure(u,i) = HH | HO&!iso(i) | OH&!iso(u) | OO&!iso(u)&!iso(i) exactly; with closed
// buildings it is a superset that exposed_slow() re-checks.  Class 0 = the pair never shares a building: its exposure
// factor (contagious3.py:230-236) is 0 on every day, so P = 0 and `P > U` can never hold -- such non-zeros are left out
// of the transposed network altogether (three quarters of the synthetic city's, most of the shipped city's).
// `shared` (optional): the ONE building the pair shares, or -1 if it shares several different ones.  With a single shared
// building b the exposure under closures is simply (class test) && open[b]; only pairs that share several buildings need
// the routine tables on a lockdown day (exposed_slow).
__device__ __forceinline__ uint32_t pair_class(const int32_t* __restrict__ ru /* [V] routine of u */,
                                               const int32_t* __restrict__ ri /* [V] routine of i */, const int V,
                                               int32_t* shared = nullptr, uint32_t* masks = nullptr)
{
    uint32_t cls = 0, m_all = 0, m_home = 0;      // slots of u that meet ANY slot of i / the home slot of i
    int32_t one = -2;                    // -2: none yet, -1: several
    for (int t = 0; t < V; ++t) {
        const int32_t bi = ri[t];
        if (bi < 0) continue;
        for (int s = 0; s < V; ++s)
            if (ru[s] == bi) {
                cls |= (s == 0) ? (t == 0 ? CLS_HH : CLS_HO) : (t == 0 ? CLS_OH : CLS_OO);
                one = (one == -2 || one == bi) ? bi : -1;
                m_all |= 1u << s;
                if (t == 0) m_home |= 1u << s;
            }
    }
    if (shared) *shared = one;
    if (masks) *masks = m_all | (m_home << 10);
    return cls;
}

// rows monotone, columns inside the halo and strictly ascending; counts, for every halo column, the non-zeros whose
// pair can ever share a building.  One warp per row.
__global__ void __launch_bounds__(256) k_validate_graph_count(const DevCtx c, const int64_t* __restrict__ rowptr,
                                                              const int32_t* __restrict__ col, int32_t* __restrict__ col_count,
                                                              int32_t* bad)
{
    const int lane = threadIdx.x & 31;
    const int64_t u = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    if (u >= c.n_loc) return;
    const int64_t e0 = rowptr[u], e1 = rowptr[u + 1];
    if (e1 < e0) { if (lane == 0) atomicOr(bad, 32); return; }
    const int V = c.V;
    int32_t ru[16];
    for (int s = 0; s < 16; ++s) ru[s] = s < V ? c.routine[(c.lo + u - c.rt_lo) * V + s] : -1;
    for (int64_t e = e0 + lane; e < e1; e += 32) {
        const int32_t i = col[e];
        if (i < c.rt_lo || i >= c.rt_lo + c.rt_n) { atomicOr(bad, 64); continue; }
        if (e > e0 && i <= col[e - 1]) atomicOr(bad, 128);
        if (pair_class(ru, c.routine + ((int64_t)i - c.rt_lo) * V, V)) atomicAdd(&col_count[i - c.rt_lo], 1);
    }
}

// exclusive prefix sum int32 counts -> int64 offsets (and a second copy used as the fill cursors); one block walks
// the array in 1024-element chunks -- load time only
__global__ void __launch_bounds__(1024) k_scan_counts(const int32_t* __restrict__ cnt, int64_t n, int64_t* __restrict__ ptr,
                                                      unsigned long long* __restrict__ cursor)
{
    __shared__ long long s_warp[32];
    __shared__ long long s_carry;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const long long v = i < n ? cnt[i] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            long long w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const long long carry = s_carry;
        const long long excl = carry + (wid ? s_warp[wid - 1] : 0) + incl - v;
        if (i < n) { ptr[i] = excl; cursor[i] = (unsigned long long)excl; }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + s_warp[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) ptr[n] = s_carry;
}

// One warp per row; each non-zero whose pair can share a building is appended to its column's transposed list (the order
// inside a list is irrelevant: the push reduces with max).
__global__ void __launch_bounds__(256) k_fill_transpose(const DevCtx c, const int64_t* __restrict__ rowptr,
                                                        const int32_t* __restrict__ col, const double* __restrict__ val,
                                                        unsigned long long* __restrict__ cursor, TEdge* __restrict__ tedge)
{
    const int lane = threadIdx.x & 31;
    const int64_t u = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    if (u >= c.n_loc) return;
    const int V = c.V;
    int32_t ru[16];
    for (int s = 0; s < 16; ++s) ru[s] = s < V ? c.routine[(c.lo + u - c.rt_lo) * V + s] : -1;
    for (int64_t e = rowptr[u] + lane; e < rowptr[u + 1]; e += 32) {
        const int32_t i = col[e];
        int32_t one = -1;
        uint32_t masks = 0;
        const uint32_t cls = pair_class(ru, c.routine + ((int64_t)i - c.rt_lo) * V, V, &one, &masks);
        if (!cls) continue;
        const unsigned long long pos = atomicAdd(&cursor[i - c.rt_lo], 1ull);
        TEdge t;
        t.colx = (uint32_t)u | (cls << CLS_SHIFT);
        t.pad = one >= 0 ? (EDGE_ONE_BLD | (uint32_t)one) : (V <= 10 ? (EDGE_MASKS | masks) : 0u);
        t.val = __dmul_rn(val[e], c.risk[u]);      // contagious3.py:232, first product (both factors are static)
        tedge[pos] = t;
    }
}

__global__ void k_keyed_uniform(uint64_t seed, uint32_t day, uint32_t stream, const uint32_t* __restrict__ a,
                                const uint32_t* __restrict__ b, int64_t n, double* __restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = keyed_uniform(seed, day, stream, a[i], b[i]);
}

}  // namespace dys
