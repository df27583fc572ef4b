// Hand-written sm_100a kernels for one day of DySTUrbD's epidemic step (contagious3.py:176-366).
//
//   k_day_start        phase A snapshot + B hospitalisation + C mortality          (contagious3.py:181-227)
//   k_contagion        phase D pairwise infection + E state transitions            (contagious3.py:230-259)
//   k_household_stats  phase F household quarantine + G integer statistics         (contagious3.py:261-366)
//   k_edge_classes     load-time: static shared-building class of every interaction_prob non-zero
//
// All three day kernels are HBM-bound integer/byte streaming passes: nothing here is a dense contraction, so
// there is no tensor-core path.  Layout and roofline reasoning: DESIGN.md.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "philox.cuh"

namespace dys {

enum : uint8_t { ST_PAD = 0, ST_S = 2, ST_I = 4, ST_QH = 6, ST_QI = 7, ST_D = 8, ST_H = 10, ST_R = 12, ST_X = 14 };

// day-start "infectious byte" of an agent: bit7 infectious today (status 2 / 3.5 / 4 at day start), bit6 isolated
// (3.5 / 4: only the home slot is visited), bits0-5 days since infection (index into the pdf table)
constexpr uint8_t IB_INF = 0x80, IB_ISO = 0x40, IB_DAYS = 0x3F;

// interaction_prob column word: bits0-26 column agent (global index), bits28-31 static shared-building class
constexpr uint32_t COL_MASK = 0x07FFFFFFu;
constexpr int CLS_SHIFT = 28;
constexpr uint32_t CLS_HH = 1u, CLS_HO = 2u, CLS_OH = 4u, CLS_OO = 8u;

constexpr int N_STATS = 64, N_SLOTS = 64, HIST0 = 32, HIST_LEN = 21;
enum {
    ST_ACTIVE = 0, ST_DAILY_INF, ST_RECOVERED, ST_QUARANTINED, ST_DAILY_QUAR, ST_HOSPITALIZED, ST_DAILY_HOSP, ST_DEAD,
    ST_DAILY_DEATHS, ST_EVER_INFECTED, ST_NEW_DIAG, ST_N_EVENTS, ST_SUS0, ST_INF0, ST_EDGES_SCANNED, ST_CANDIDATES,
    ST_EXPOSED, ST_HITS, ST_CHANGED
};

struct DevCtx {
    int64_t n_loc, n_pad, n_glob, lo;    // owned agents, padded count (multiple of 1024), global agents, first owned
    int64_t rt_lo, rt_n;                 // routine[] covers global agents [rt_lo, rt_lo + rt_n)
    int64_t H, B, S;
    int32_t V;
    int32_t recover, hospital_recover, diagnosis, quarantine, adm_min, adm_max, death_min;
    double norm_factor, compliance;
    uint64_t seed;
    // mutable state, owned agents
    uint8_t* status;
    int16_t *t_inf, *t_q, *t_adm;
    int32_t* src;
    // static per owned agent
    const int32_t *hh, *home, *sa;
    const double *risk, *p_adm, *p_death;
    const int32_t* routine;              // [rt_n * V] building index or -1 (owned agents + halo: neighbours' rows are read)
    const int64_t* rowptr;               // [n_pad + 1]
    const uint32_t* colx;                // [nnz (+pad)] column | class << 28
    const double* val;                   // [nnz]
    const double* pdf;                   // [64]
    uint8_t* ib;                         // [n_glob padded] day-start infectious bytes of ALL agents
    const uint8_t* open;                 // [B]
    uint8_t *hh_flag, *hh_qflag;         // [H] today's diagnosed households / quirk households
    const uint8_t* hh_always;            // [H] or null
    const int64_t* trig_ptr;             // [n_pad + 1] or null
    const int32_t* trig_hh;
    int32_t* day_flags;                  // [0] any new diagnosis today, [1] events claimed today, [2] closed buildings
    unsigned long long* part;            // [N_STATS][N_SLOTS] spread partial sums
    int64_t* stats;                      // [N_STATS]
    unsigned int* ticket;
    int32_t* sa_hist;                    // [S * 21] or null
    int32_t* bcount;                     // [B] or null
    int32_t *ev_agent, *ev_src;          // [n_pad] or null
    int32_t* io_mat;                     // [S * S] or null
    const double *u_adm, *u_death, *u_pair;   // injected draws or null
    double* pair_prob;                   // [n_glob^2] or null (tests)
};

__device__ __forceinline__ double draw(const double* inj, int64_t idx, uint64_t seed, uint32_t day, uint32_t stream,
                                       uint32_t a, uint32_t b)
{
    return inj ? __ldg(inj + idx) : keyed_uniform(seed, day, stream, a, b);
}

// block-level: add per-thread counters into the spread partial-sum slots (one atomic per counter per block)
template <int NC>
__device__ __forceinline__ void block_accumulate(const DevCtx& c, const int (&idx)[NC], unsigned int (&v)[NC])
{
    __shared__ unsigned int s_acc[NC];
    if (threadIdx.x < NC) s_acc[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        unsigned int x = v[k];
#pragma unroll
        for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0 && x) atomicAdd(&s_acc[k], x);
    }
    __syncthreads();
    if (threadIdx.x < NC && s_acc[threadIdx.x])
        atomicAdd(&c.part[(size_t)idx[threadIdx.x] * N_SLOTS + (blockIdx.x & (N_SLOTS - 1))],
                  (unsigned long long)s_acc[threadIdx.x]);
}

template <typename T>
__device__ __forceinline__ void grid_clear(T* p, int64_t n)
{
    if (!p) return;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = T(0);
}

// ---------------------------------------------------------------------------------------------------------------
// k_day_start: 4 agents per thread, vector loads.  Writes the day-start snapshot byte of every owned agent and
// applies the two per-agent draws.  Also clears the per-day scratch (nothing else touches it during this kernel).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_day_start(const DevCtx c, const int day)
{
    grid_clear(c.hh_flag, c.H);
    grid_clear(c.hh_qflag, c.H);
    grid_clear(c.sa_hist, c.S * HIST_LEN);
    grid_clear(c.bcount, c.B);
    grid_clear(c.io_mat, c.io_mat ? c.S * c.S : 0);
    if (blockIdx.x == 0 && threadIdx.x < 2) c.day_flags[threadIdx.x] = 0;

    const int64_t a0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    unsigned int cnt[4] = {0, 0, 0, 0};   // sus0, inf0, admissions, deaths
    if (a0 < c.n_pad) {
        uchar4 sv = *reinterpret_cast<const uchar4*>(c.status + a0);
        uint8_t s[4] = {sv.x, sv.y, sv.z, sv.w};
        bool anyinf = false, anyh = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            anyinf |= (s[j] == ST_I) | (s[j] == ST_QI) | (s[j] == ST_D);
            anyh |= (s[j] == ST_H);
        }
        short4 ti4 = make_short4(0, 0, 0, 0), ta4 = make_short4(0, 0, 0, 0);
        if (anyinf) ti4 = *reinterpret_cast<const short4*>(c.t_inf + a0);
        if (anyh) ta4 = *reinterpret_cast<const short4*>(c.t_adm + a0);
        const int16_t ti[4] = {ti4.x, ti4.y, ti4.z, ti4.w}, ta[4] = {ta4.x, ta4.y, ta4.z, ta4.w};
        uint8_t ibv[4];
        bool changed = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t a = a0 + j;
            const uint8_t st = s[j];
            ibv[j] = 0;
            cnt[0] += (st == ST_S) | (st == ST_QH);
            if ((st == ST_I) | (st == ST_QI) | (st == ST_D)) {
                ++cnt[1];
                int n = day - ti[j];
                n = n < 0 ? 0 : (n > 63 ? 63 : n);
                ibv[j] = IB_INF | (st != ST_I ? IB_ISO : 0) | (uint8_t)n;
                if (ti[j] >= c.adm_min && ti[j] <= c.adm_max) {        // contagious3.py:205 (absolute infection day)
                    const double u = draw(c.u_adm, c.lo + a, c.seed, day, STREAM_ADMISSION, (uint32_t)(c.lo + a), 0);
                    if (c.p_adm[a] > u) {                               // :213
                        s[j] = ST_H; c.t_adm[a] = (int16_t)day; ++cnt[2]; changed = true;
                    }
                }
            } else if (st == ST_H) {
                if (ta[j] >= c.death_min) {                             // :219
                    const double u = draw(c.u_death, c.lo + a, c.seed, day, STREAM_DEATH, (uint32_t)(c.lo + a), 0);
                    if (c.p_death[a] > u) {                             // :225
                        s[j] = ST_X; c.t_adm[a] = 0; ++cnt[3]; changed = true;   // :259 resets col 24 for the dead
                    }
                }
            }
        }
        *reinterpret_cast<uchar4*>(c.ib + c.lo + a0) = make_uchar4(ibv[0], ibv[1], ibv[2], ibv[3]);
        if (changed) *reinterpret_cast<uchar4*>(c.status + a0) = make_uchar4(s[0], s[1], s[2], s[3]);
    }
    const int idx[4] = {ST_SUS0, ST_INF0, ST_DAILY_HOSP, ST_DAILY_DEATHS};
    block_accumulate<4>(c, idx, cnt);
}

// exact exposure test for days with closed buildings (contagious3.py:181-189, :230-231): a shared building must be
// open, and an isolated agent only keeps slot 0
__device__ __noinline__ bool exposed_slow(const int32_t* __restrict__ routine, const uint8_t* __restrict__ open, int V,
                                          int64_t ug, int64_t ig, bool iso_u, bool iso_i)
{
    const int32_t* ru = routine + ug * V;
    const int32_t* ri = routine + ig * V;
    for (int s = 0; s < V; ++s) {
        const int32_t bu = __ldg(ru + s);
        if (bu < 0 || (s > 0 && iso_u) || !__ldg(open + bu)) continue;
        for (int t = 0; t < V; ++t) {
            if (t > 0 && iso_i) break;
            if (__ldg(ri + t) == bu) return true;
        }
    }
    return false;
}

// ---------------------------------------------------------------------------------------------------------------
// k_contagion: a block owns a tile of 256 consecutive agents = one CONTIGUOUS run of interaction_prob column words.
// The run is streamed in lock step with aligned 128-bit loads (two chunks = 8 edges per thread in flight), each
// column's day-start infectious byte is gathered (L1/L2 resident: neighbours live in the same community), and only
// for an infectious column is the row looked up (binary search over the tile's row offsets staged in shared
// memory).  The infector of a newly infected agent is the LARGEST infecting row index (contagious3.py:248) = a max,
// kept per row with shared-memory atomicMax (order independent, no global atomics).
// ---------------------------------------------------------------------------------------------------------------
constexpr int TILE = 256;

__global__ void __launch_bounds__(TILE, 4) k_contagion(const DevCtx c, const int day)
{
    __shared__ int s_rp[TILE + 1];
    __shared__ int s_best[TILE];
    __shared__ uint8_t s_st[TILE];
    __shared__ double s_pdf[64];
    __shared__ int s_nev, s_evbase;
    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t tile0 = (int64_t)blockIdx.x * TILE;     // n_pad is a multiple of 1024
    const int64_t a = tile0 + tid;
    const int64_t e0 = c.rowptr[tile0];
    const uint8_t st0 = c.status[a];
    const bool sus = (st0 == ST_S) | (st0 == ST_QH);
    s_rp[tid + 1] = (int)(c.rowptr[a + 1] - e0);
    s_st[tid] = st0;
    s_best[tid] = -1;
    if (tid == 0) { s_rp[0] = 0; s_nev = 0; }
    if (tid < 64) s_pdf[tid] = c.pdf[tid];
    const int any_sus = __syncthreads_or(sus);
    const int T = s_rp[TILE];
    unsigned int n_scan = sus ? (unsigned)(s_rp[tid + 1] - s_rp[tid]) : 0u, n_cand = 0, n_exp = 0, n_hit = 0;

    if (any_sus && T > 0) {
        const bool any_closed = c.day_flags[2] != 0;
        const int64_t eend = e0 + T;
        auto candidate = [&](const int64_t e, const uint32_t cw, const uint8_t ibv) {
            const int rel = (int)(e - e0);
            int lo = 0, hi = TILE;                         // s_rp[lo] <= rel < s_rp[hi]
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_rp[mid] <= rel) lo = mid; else hi = mid;
            }
            const uint8_t st_r = s_st[lo];
            if (st_r != ST_S && st_r != ST_QH) return;
            ++n_cand;
            const int64_t ul = tile0 + lo, ug = c.lo + ul;
            const uint32_t ig = cw & COL_MASK, cls = cw >> CLS_SHIFT;
            const bool iso_u = (st_r == ST_QH), iso_i = ibv & IB_ISO;
            bool ex = (cls & CLS_HH) | ((cls & CLS_HO) && !iso_i) | ((cls & CLS_OH) && !iso_u) |
                      ((cls & CLS_OO) && !iso_u && !iso_i);
            if (ex && any_closed) ex = exposed_slow(c.routine, c.open, c.V, ug - c.rt_lo, (int64_t)ig - c.rt_lo, iso_u, iso_i);
            if (!ex) return;
            ++n_exp;
            // contagious3.py:232-238, same association, separately rounded (no FMA can form: no adds)
            double p = __dmul_rn(__ldg(c.val + e), __ldg(c.risk + ul));
            p = __dmul_rn(p, s_pdf[ibv & IB_DAYS]);
            p = __dmul_rn(p, c.norm_factor);
            if (c.pair_prob) c.pair_prob[(size_t)ug * c.n_glob + ig] = p;
            const double u = draw(c.u_pair, ug * c.n_glob + ig, c.seed, day, STREAM_INFECTION, (uint32_t)ug, ig);
            if (p > u) { ++n_hit; atomicMax(&s_best[lo], (int)ig); }        // :240, :248
        };
        for (int64_t base = (e0 & ~3LL) + 4 * tid; base < eend; base += 8 * TILE) {
            const int64_t base1 = base + 4 * TILE;
            const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(c.colx + base));
            uint4 v1 = make_uint4(0, 0, 0, 0);
            if (base1 < eend) v1 = __ldg(reinterpret_cast<const uint4*>(c.colx + base1));
            const uint32_t cw[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            uint8_t ibv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int64_t e = (j < 4 ? base : base1 - 4) + j;
                ibv[j] = (e >= e0 && e < eend) ? __ldg(c.ib + (cw[j] & COL_MASK)) : (uint8_t)0;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (ibv[j] & IB_INF) candidate((j < 4 ? base : base1 - 4) + j, cw[j], ibv[j]);
        }
    }
    __syncthreads();

    // ---- per-agent: apply infection (contagious3.py:241-248), then the ordered rules of :250-259 ----
    const int best = s_best[tid];
    uint8_t s = st0;
    int tinf = 0, tq = 0;
    bool w_tinf = false, w_tq = false, w_tadm = false;
    if (s == ST_I || s == ST_QI || s == ST_D || s == ST_H) tinf = c.t_inf[a];
    if (s == ST_QH || s == ST_QI) tq = c.t_q[a];
    const bool infected_today = best >= 0;
    if (infected_today) {
        s = (st0 == ST_S) ? ST_I : ST_QI;
        tinf = day; w_tinf = true;
        c.src[a] = best;
    }
    const int n_inf = day - tinf, n_q = day - tq;
    if (s == ST_QH && n_q == c.quarantine) s = ST_S;                                   // :250
    if (s == ST_QI && n_q == c.quarantine) s = ST_I;                                   // :252
    if (s == ST_I && n_inf == c.diagnosis) { tq = day; w_tq = true; }                  // :253
    if ((s == ST_I || s == ST_QI) && n_inf == c.diagnosis) s = ST_D;                   // :254
    if (s == ST_D && n_inf == c.recover) s = ST_R;                                     // :255
    if (s == ST_H && n_inf == c.hospital_recover) { s = ST_R; w_tadm = true; }         // :256, :259
    const bool new_diag = (s == ST_D && n_inf == c.diagnosis);                         // :261
    if (w_tinf) c.t_inf[a] = (int16_t)tinf;
    if (w_tq) c.t_q[a] = (int16_t)tq;
    if (w_tadm) c.t_adm[a] = 0;
    if (s != st0) c.status[a] = s;
    if (new_diag) {
        c.hh_flag[c.hh[a]] = 1;                       // same-value stores: race-free in effect
        if (c.trig_ptr)
            for (int64_t t = c.trig_ptr[a]; t < c.trig_ptr[a + 1]; ++t) c.hh_qflag[c.trig_hh[t]] = 1;
        c.day_flags[0] = 1;
    }

    // ---- today's infection events, block-compacted (order across blocks is arbitrary; hosts sort) ----
    if (c.ev_agent || c.io_mat) {
        const unsigned evm = __ballot_sync(0xffffffffu, infected_today);
        int woff = 0;
        if (lane == 0 && evm) woff = atomicAdd(&s_nev, __popc(evm));
        woff = __shfl_sync(0xffffffffu, woff, 0);
        __syncthreads();
        if (threadIdx.x == 0 && s_nev) s_evbase = atomicAdd(&c.day_flags[1], s_nev);
        __syncthreads();
        if (infected_today) {
            if (c.ev_agent) {
                const int slot = s_evbase + woff + __popc(evm & ((1u << lane) - 1u));
                c.ev_agent[slot] = (int)(c.lo + a);
                c.ev_src[slot] = best;
            }
            if (c.io_mat) {
                // contagious3.py:356-366: row = area of the infected, column = area of the infector
                const int sa_u = c.sa[a], sa_i = c.sa[best - c.lo];
                if (sa_u >= 0 && sa_i >= 0) atomicAdd(&c.io_mat[(size_t)sa_u * c.S + sa_i], 1);
            }
        }
    }
    unsigned int cnt[7] = {n_scan, n_cand, n_exp, n_hit, (unsigned)infected_today, (unsigned)new_diag,
                           (unsigned)(s != st0 || w_tq)};
    const int idx[7] = {ST_EDGES_SCANNED, ST_CANDIDATES, ST_EXPOSED, ST_HITS, ST_N_EVENTS, ST_NEW_DIAG, ST_CHANGED};
    block_accumulate<7>(c, idx, cnt);
}

// ---------------------------------------------------------------------------------------------------------------
// k_household_stats: phase F on the household flags raised by k_contagion, then the integer statistics of the
// end-of-day state.  The last block to finish folds the spread partial sums into the stats block.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_household_stats(const DevCtx c, const int day)
{
    __shared__ unsigned int s_hist[HIST_LEN];
    __shared__ bool s_last;
    if (threadIdx.x < HIST_LEN) s_hist[threadIdx.x] = 0;
    __syncthreads();
    const bool any_nd = c.day_flags[0] != 0;
    const int64_t a0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    // active, daily_inf, recovered, quarantined, daily_quar, hosp, dead, ever, changed
    unsigned int cnt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (a0 < c.n_pad) {
        const uchar4 sv = *reinterpret_cast<const uchar4*>(c.status + a0);
        uint8_t s[4] = {sv.x, sv.y, sv.z, sv.w};
        bool changed = false, any_ever = false, any_q = false;
        bool newq[4] = {false, false, false, false};
        if (any_nd) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t a = a0 + j;
                if (s[j] != ST_S && s[j] != ST_I) continue;
                const int h = c.hh[a];
                bool hit = c.hh_flag[h];                                              // contagious3.py:263
                if (s[j] == ST_I && !hit)                                             // :267 (whole-row isin)
                    hit = c.hh_qflag[h] || (c.hh_always && c.hh_always[h]);
                if (hit && c.compliance < 1.0)
                    hit = keyed_uniform(c.seed, day, STREAM_COMPLIANCE, (uint32_t)(c.lo + a), 0) < c.compliance;
                if (hit) {
                    s[j] = (s[j] == ST_S) ? ST_QH : ST_QI;                            // :265, :269
                    c.t_q[a] = (int16_t)day;                                          // :266, :270
                    newq[j] = true; changed = true; ++cnt[8];
                }
            }
            if (changed) *reinterpret_cast<uchar4*>(c.status + a0) = make_uchar4(s[0], s[1], s[2], s[3]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            any_ever |= !(s[j] == ST_S || s[j] == ST_QH || s[j] == ST_PAD);
            any_q |= (s[j] == ST_QH || s[j] == ST_QI || s[j] == ST_D);
        }
        short4 ti4 = make_short4(0, 0, 0, 0), tq4 = make_short4(0, 0, 0, 0);
        if (any_ever) ti4 = *reinterpret_cast<const short4*>(c.t_inf + a0);
        if (any_q) tq4 = *reinterpret_cast<const short4*>(c.t_q + a0);
        const int16_t ti[4] = {ti4.x, ti4.y, ti4.z, ti4.w}, tq[4] = {tq4.x, tq4.y, tq4.z, tq4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t a = a0 + j;
            const uint8_t st = s[j];
            if (st == ST_PAD) continue;
            const bool ever = !(st == ST_S || st == ST_QH);
            const bool quar = (st == ST_QH || st == ST_QI || st == ST_D);
            cnt[0] += (st == ST_I || st == ST_QI || st == ST_D || st == ST_H);
            cnt[2] += (st == ST_R);
            cnt[3] += quar;
            cnt[4] += quar && (newq[j] || tq[j] == day);
            cnt[5] += (st == ST_H);
            cnt[6] += (st == ST_X);
            cnt[7] += ever;
            if (ever) {
                const int k = day - ti[j];
                cnt[1] += (k == 0);
                if (k >= 0 && k < HIST_LEN) {
                    atomicAdd(&s_hist[k], 1u);
                    if (c.sa_hist) { const int sa = c.sa[a]; if (sa >= 0) atomicAdd(&c.sa_hist[(size_t)sa * HIST_LEN + k], 1); }
                }
            }
            if (c.bcount && (st == ST_I || st == ST_QI || st == ST_D)) atomicAdd(&c.bcount[c.home[a]], 1);
        }
    }
    const int idx[9] = {ST_ACTIVE, ST_DAILY_INF, ST_RECOVERED, ST_QUARANTINED, ST_DAILY_QUAR, ST_HOSPITALIZED, ST_DEAD,
                        ST_EVER_INFECTED, ST_CHANGED};
    block_accumulate<9>(c, idx, cnt);
    if (threadIdx.x < HIST_LEN && s_hist[threadIdx.x])
        atomicAdd(&c.part[(size_t)(HIST0 + threadIdx.x) * N_SLOTS + (blockIdx.x & (N_SLOTS - 1))],
                  (unsigned long long)s_hist[threadIdx.x]);

    // last block folds the partial sums (threadfence reduction) and re-zeroes them for the next day
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(c.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (s_last) {
        __threadfence();
        // part[counter][slot]: a warp folds one counter with two coalesced loads and a shuffle tree
        const int lane = threadIdx.x & 31;
        for (int k = threadIdx.x >> 5; k < N_STATS; k += 8) {
            unsigned long long* p = c.part + (size_t)k * N_SLOTS;
            unsigned long long t = __ldcg(p + lane) + __ldcg(p + lane + 32);
            p[lane] = 0ull;
            p[lane + 32] = 0ull;
#pragma unroll
            for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (lane == 0) c.stats[k] = (int64_t)t;
        }
        if (threadIdx.x == 0) *c.ticket = 0u;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k_edge_classes (load time): routines are static (computed once per simulation, contagious3.py:139-170), so
// WHICH slots two agents share is static too.  Four bits per interaction_prob non-zero (u = row, i = column):
//   HH  home(u) == home(i)            HO  home(u) is a non-home stop of i
//   OH  home(i) is a non-home stop of u     OO  they share a non-home stop
// With every building open, exposure(u,i) = HH | HO&!iso(i) | OH&!iso(u) | OO&!iso(u)&!iso(i) exactly; with closed
// buildings it is a superset that exposed_slow() re-checks.  One warp per row.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_edge_classes(const DevCtx c, const int32_t* __restrict__ col, uint32_t* __restrict__ colx)
{
    const int lane = threadIdx.x & 31;
    const int64_t u = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    if (u >= c.n_loc) return;
    const int V = c.V;
    int32_t ru[16];
    for (int s = 0; s < 16; ++s) ru[s] = s < V ? c.routine[(c.lo + u - c.rt_lo) * V + s] : -1;
    for (int64_t e = c.rowptr[u] + lane; e < c.rowptr[u + 1]; e += 32) {
        const int32_t i = col[e];
        const int32_t* ri = c.routine + ((int64_t)i - c.rt_lo) * V;
        uint32_t cls = 0;
        for (int t = 0; t < V; ++t) {
            const int32_t bi = ri[t];
            if (bi < 0) continue;
            for (int s = 0; s < V; ++s) {
                if (ru[s] != bi) continue;
                cls |= (s == 0) ? (t == 0 ? CLS_HH : CLS_HO) : (t == 0 ? CLS_OH : CLS_OO);
            }
        }
        colx[e] = (uint32_t)i | (cls << CLS_SHIFT);
    }
}

// int32 -> int16 day columns with range check (days are < 32768)
__global__ void k_narrow_days(const int32_t* __restrict__ in, int16_t* __restrict__ out, int64_t n, int32_t* bad)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t v = in[i];
    if (v < -32768 || v > 32767) atomicOr(bad, 256);
    out[i] = (int16_t)v;
}

__global__ void k_widen_days(const int16_t* __restrict__ in, int32_t* __restrict__ out, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

// input validation: every index inside its range; t_adm == 0 wherever status is in {1, 2, 6, 7}
// (contagious3.py:259 keeps that invariant, and the kernels rely on it to skip the column)
__global__ void k_validate_population(const DevCtx c, int32_t* bad)
{
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= c.n_loc) return;
    const uint8_t s = c.status[a];
    const bool known = s == ST_S || s == ST_I || s == ST_QH || s == ST_QI || s == ST_D || s == ST_H || s == ST_R || s == ST_X;
    if (!known) atomicOr(bad, 1);
    if (c.hh[a] < 0 || c.hh[a] >= c.H || c.home[a] < 0 || c.home[a] >= c.B || c.sa[a] < -1 || c.sa[a] >= c.S) atomicOr(bad, 2);
    if ((s == ST_S || s == ST_I || s == ST_R || s == ST_X) && c.t_adm[a] != 0) atomicOr(bad, 4);
    if (c.src[a] < -1 || c.src[a] >= c.n_glob) atomicOr(bad, 8);
}

__global__ void k_validate_routine(const int32_t* __restrict__ routine, int64_t n, int64_t B, int32_t* bad)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && (routine[i] < -1 || routine[i] >= B)) atomicOr(bad, 16);
}

__global__ void k_validate_graph(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n_loc,
                                 int64_t col_lo, int64_t col_hi, int32_t* bad)
{
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_loc) return;
    if (rowptr[u + 1] < rowptr[u]) { atomicOr(bad, 32); return; }
    int32_t prev = -1;
    for (int64_t e = rowptr[u]; e < rowptr[u + 1]; ++e) {
        const int32_t i = col[e];
        if (i < col_lo || i >= col_hi) atomicOr(bad, 64);
        if (i <= prev) atomicOr(bad, 128);      // columns must ascend: "last writer" == largest index
        prev = i;
    }
}

__global__ void k_keyed_uniform(uint64_t seed, uint32_t day, uint32_t stream, const uint32_t* __restrict__ a,
                                const uint32_t* __restrict__ b, int64_t n, double* __restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = keyed_uniform(seed, day, stream, a[i], b[i]);
}

__global__ void k_count_closed(const uint8_t* __restrict__ open, int64_t B, int32_t* out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B && !open[i]) atomicAdd(out, 1);
}

}  // namespace dys
