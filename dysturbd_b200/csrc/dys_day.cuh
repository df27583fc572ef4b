// The two kernels of a simulated day (contagious3.py:176-366), plus the stand-alone snapshot.
//
//   k_push   phase D, pairwise infection (contagious3.py:230-248), as a FRONTIER PUSH: only today's infectious agents
//            do work.  For infectious i the transposed interaction_prob row lists every owned u with ip[u,i] != 0; a
//            susceptible u that shares a building with i today draws one Bernoulli; a success does
//            atomicMax(src[u], i) -- the reference's "last writer wins" is the LARGEST infecting row index (:248), an
//            order-independent max, and src[u] is -1 for every susceptible, so the infector array itself is the mailbox.
//   k_agent  everything that is per agent, one streaming pass: B hospitalisation, C mortality, applying D, the ordered
//            rules of E, F household quarantine, G the integer statistics -- and tomorrow's phase A (snapshot byte and
//            the households that WILL see a diagnosis tomorrow, which is a pure function of an agent's own state and
//            its keyed admission draw), so a day is two launches.
//   k_snapshot  phase A on its own: first day after a load / dys_set_state, and every day in injected-draw mode.
//
// Both day kernels are HBM/L2-bound integer and byte work; there is no dense contraction, hence no tensor-core path.
#pragma once
#include "dys_common.cuh"

namespace dys {

// Will agent a (day-start status st, infection day tinf) be among `new_diagnosed_agents` (contagious3.py:261) on
// `day`?  Yes iff it is infectious, day - tinf == diagnosis (rule :254 fires) and phase B does not hospitalise it first.
__device__ __forceinline__ bool diagnosed_on(const DevCtx& c, const double* u_adm, int64_t a, uint32_t st, int tinf, int day)
{
    if (!in_set(M_INF, st) || day - tinf != c.diagnosis) return false;
    if (tinf >= c.adm_min && tinf <= c.adm_max) {
        const double u = draw(u_adm, c.lo + a, c.seed, day, STREAM_ADMISSION, (uint32_t)(c.lo + a), 0);
        if (__ldg(c.p_adm + a) > u) return false;
    }
    return true;
}

__device__ __forceinline__ void raise_household(const DevCtx& c, int64_t a, uint8_t* hh_flag, uint8_t* hh_qflag, int32_t* df)
{
    hh_flag[c.hh[a]] = 1;                             // same-value byte stores: race-free in effect
    if (c.trig_ptr)
        for (int64_t t = c.trig_ptr[a]; t < c.trig_ptr[a + 1]; ++t) hh_qflag[c.trig_hh[t]] = 1;
    df[DF_ANY_ND] = 1;
}

// ---------------------------------------------------------------------------------------------------------------
// k_push.  Persistent warps walk groups of 32 consecutive halo agents.  The infectious lanes' transposed rows are
// flattened (warp prefix sum of their lengths) so that all 32 lanes work on consecutive entries regardless of how the
// infectious agents are spread; each entry costs one coalesced 4-byte read and one 1-byte status gather.
// ---------------------------------------------------------------------------------------------------------------
constexpr int PUSH_WARPS = 8;

__global__ void __launch_bounds__(PUSH_WARPS * 32) k_push(const DevCtx c, const int day)
{
    __shared__ double s_pdf[64];
    __shared__ unsigned int s_acc[8];
    __shared__ int s_off[PUSH_WARPS][36];
    __shared__ int64_t s_t0[PUSH_WARPS][32];
    __shared__ uint8_t s_ib[PUSH_WARPS][32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    // scratch that k_agent fills today, and the flags it raises for tomorrow
    grid_clear(c.sa_hist, c.S * HIST_LEN);
    grid_clear(c.bcount, c.B);
    grid_clear(c.io_mat, c.io_mat ? c.S * c.S : 0);
    grid_clear16(c.hh_flag_next, c.H);
    grid_clear16(c.hh_qflag_next, c.H);
    if (blockIdx.x == 0 && tid < DF_LEN) c.df_next[tid] = 0;

    if (tid < 64) s_pdf[tid] = c.pdf[tid];
    if (tid < 8) s_acc[tid] = 0;
    __syncthreads();
    const bool any_closed = c.n_closed != 0;
    unsigned int cnt[4] = {0, 0, 0, 0};   // entries scanned, candidates, exposed, hits

    const int64_t n_groups = (c.rt_n + 31) >> 5;
    for (int64_t g = (int64_t)blockIdx.x * PUSH_WARPS + wid; g < n_groups; g += (int64_t)gridDim.x * PUSH_WARPS) {
        const int64_t il = g * 32 + lane;                 // halo-relative index of this lane's agent
        const uint32_t ibv = il < c.rt_n ? (uint32_t)c.ib[c.rt_lo + il] : 0u;
        const bool inf = ibv & IB_INF;
        if (!__ballot_sync(0xffffffffu, inf)) continue;
        int64_t t0 = 0;
        int len = 0;
        if (inf) { t0 = c.tptr[il]; len = (int)(c.tptr[il + 1] - t0); }
        int incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        s_off[wid][lane] = incl - len;
        if (lane == 31) s_off[wid][32] = total;
        s_t0[wid][lane] = t0;
        s_ib[wid][lane] = (uint8_t)ibv;
        __syncwarp();
        cnt[0] += len;
        for (int f0 = 0; f0 < total; f0 += 32) {
            const int f = f0 + lane;
            if (f >= total) continue;
            int r = 0, hi = 32;                            // s_off[r] <= f < s_off[hi]; the last such r is the row
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const int mid = (r + hi) >> 1;
                if (s_off[wid][mid] <= f) r = mid; else hi = mid;
            }
            const int64_t e = s_t0[wid][r] + (f - s_off[wid][r]);
            const uint32_t cw = __ldg(c.tcolx + e);
            const uint32_t ul = cw & COL_MASK, cls = cw >> CLS_SHIFT;
            const uint32_t st_u = c.status[ul];           // day-start status: k_agent has not run yet today
            if (!in_set(M_SUS, st_u)) continue;
            ++cnt[1];
            const uint32_t ib_i = s_ib[wid][r];
            const bool iso_u = (st_u == ST_QH), iso_i = ib_i & IB_ISO;
            const int64_t ig = c.rt_lo + g * 32 + r, ug = c.lo + ul;
            bool ex = (cls & CLS_HH) | ((cls & CLS_HO) && !iso_i) | ((cls & CLS_OH) && !iso_u) |
                      ((cls & CLS_OO) && !iso_u && !iso_i);
            if (ex && any_closed) ex = exposed_slow(c.routine, c.open, c.V, ug - c.rt_lo, ig - c.rt_lo, iso_u, iso_i);
            if (!ex) continue;
            ++cnt[2];
            // contagious3.py:232-238: ((ip[u,i] * risk[u]) * pdf(days since i's infection)) * norm_factor, each product
            // rounded separately (no add follows, so no FMA can form)
            double p = __dmul_rn(__ldg(c.tval + e), __ldg(c.risk + ul));
            p = __dmul_rn(p, s_pdf[ib_i & IB_DAYS]);
            p = __dmul_rn(p, c.norm_factor);
            if (c.pair_prob) c.pair_prob[(size_t)ug * c.n_glob + ig] = p;
            const double u = draw(c.u_pair, ug * c.n_glob + ig, c.seed, day, STREAM_INFECTION, (uint32_t)ug, (uint32_t)ig);
            if (p > u) { ++cnt[3]; atomicMax(&c.src[ul], (int)ig); }           // :240, :248
        }
        __syncwarp();
    }
    const int idx[4] = {ST_EDGES_SCANNED, ST_CANDIDATES, ST_EXPOSED, ST_HITS};
    __syncthreads();
    block_accumulate<4>(c, idx, cnt, s_acc);
}

// ---------------------------------------------------------------------------------------------------------------
// k_agent.  4 agents per thread, vector loads, one pass; columns an agent's status cannot need are not loaded, and an
// agent to whom nothing can happen today (susceptible, not infected, no household diagnosed anywhere; recovered;
// dead) only feeds the counters.  Compartment counts are reduced as packed status histograms (two REDUX per warp);
// rare events (admissions, deaths, diagnoses, infections, ...) go straight to shared-memory counters.
// ---------------------------------------------------------------------------------------------------------------
// packed 8-bit histogram field of a status code: QI S I QH | D H R X   (4 agents x 32 lanes = 128 < 256 per field)
__device__ __forceinline__ void hist_add(uint32_t (&w)[2], uint32_t st)
{
    const uint32_t idx = (st == ST_QI) ? 0u : (st >> 1);
    w[idx >> 2] += 1u << (8u * (idx & 3u));
}
constexpr uint32_t M_F = M_SUS | sbit(ST_I) | sbit(ST_QI);      // can be in status 1 or 2 when F runs
enum { EV_ADM = 0, EV_DEATH, EV_DAILY_INF, EV_DAILY_QUAR, EV_NEW_DIAG, EV_EVENTS, EV_CHANGED, EV_LEN };
// internal partial-sum rows (folded into the public stats by the last block)
constexpr int RAW_END = 53;      // [53..60] end-of-day status counts in hist_add order
constexpr int RAW_SUS0 = 61, RAW_INF0 = 62;
constexpr int ST_COUNT0 = 20;    // public: stats[20..27] = end-of-day agents in S, I, QH, QI, D, H, R, X

template <int VARIANT>   // 0 = product; 1, 2 = tools/kbench.cu ablations (no fold / no reduction at all)
__device__ __forceinline__ void agent_body(const DevCtx& c, const int day)
{
    __shared__ unsigned int s_hist[HIST_LEN];
    __shared__ unsigned int s_ev[EV_LEN];
    __shared__ unsigned int s_cnt[10];            // 8 status counts, sus0, inf0
    __shared__ long long s_fold[N_STATS];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < HIST_LEN) s_hist[threadIdx.x] = 0;
    if (threadIdx.x < EV_LEN) s_ev[threadIdx.x] = 0;
    if (threadIdx.x < 10) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const bool any_nd = c.df[DF_ANY_ND] != 0;
    const int64_t a0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    uint32_t h_end[2] = {0, 0}, h_start = 0;      // h_start: sus0 | inf0 << 8
    uint32_t ev_mask = 0;
    int4 src4 = make_int4(-1, -1, -1, -1);
    if (a0 < c.n_pad) {
        const uchar4 sv = *reinterpret_cast<const uchar4*>(c.status + a0);
        const uint32_t st0[4] = {sv.x, sv.y, sv.z, sv.w};
        uint32_t m_sus = 0, m_need_ti = 0, m_iso = 0, m_h = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            m_sus |= in_set(M_SUS, st0[j]) << j;
            m_need_ti |= in_set(M_EVER, st0[j]) << j;
            m_iso |= in_set(M_ISO, st0[j]) << j;
            m_h |= (st0[j] == ST_H) << j;
        }
        short4 ti4 = make_short4(0, 0, 0, 0), tq4 = make_short4(0, 0, 0, 0), ta4 = make_short4(0, 0, 0, 0);
        if (m_need_ti) ti4 = *reinterpret_cast<const short4*>(c.t_inf + a0);
        if (m_iso) tq4 = *reinterpret_cast<const short4*>(c.t_q + a0);
        if (m_h) ta4 = *reinterpret_cast<const short4*>(c.t_adm + a0);
        if (m_sus) src4 = *reinterpret_cast<const int4*>(c.src + a0);        // the mailbox k_push wrote into
        int ti[4] = {ti4.x, ti4.y, ti4.z, ti4.w}, tq[4] = {tq4.x, tq4.y, tq4.z, tq4.w}, ta[4] = {ta4.x, ta4.y, ta4.z, ta4.w};
        const int srcv[4] = {src4.x, src4.y, src4.z, src4.w};
        uint32_t s[4];
        bool w_st = false;
        uint32_t w_ti = 0, w_tq = 0, w_ta = 0;     // per-agent write masks (columns are only loaded when needed)
        uint32_t ibw = 0;                          // tomorrow's four snapshot bytes
        // F needs the household flags of agents that can be in status 1 or 2 after E: fetch all four up front so
        // the gathers overlap instead of serialising behind each agent's rules
        uint32_t hflag[4] = {0, 0, 0, 0};          // bit0 household diagnosed today, bit1 quirk household
        if (any_nd) {
            uint32_t m_f = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) m_f |= in_set(M_F, st0[j]) << j;
            if (m_f) {
                const int4 hh4 = __ldg(reinterpret_cast<const int4*>(c.hh + a0));
                const int hv[4] = {hh4.x, hh4.y, hh4.z, hh4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if ((m_f >> j) & 1u) {
                        uint32_t f = __ldg(c.hh_flag + hv[j]);
                        f |= (uint32_t)(__ldg(c.hh_qflag + hv[j]) | (c.hh_always ? __ldg(c.hh_always + hv[j]) : 0)) << 1;
                        hflag[j] = f;
                    }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t st = st0[j];
            s[j] = st;
            if (st == ST_PAD) continue;
            const bool sus = (m_sus >> j) & 1u;
            h_start += sus ? 1u : (in_set(M_INF, st) ? 0x100u : 0u);
            // nothing can happen to: a susceptible agent nobody infected whose household is not flagged; recovered; dead
            const bool idle = (st == ST_S) ? (srcv[j] < 0 && !any_nd) : (st == ST_R || st == ST_X);
            uint32_t x = st;
            if (!idle) {
                const int64_t a = a0 + j;
                // ---- B, C (contagious3.py:205-227) on the day-start status; D's result from the mailbox ----
                if (in_set(M_INF, st)) {
                    if (ti[j] >= c.adm_min && ti[j] <= c.adm_max) {              // :205 window on the ABSOLUTE infection day
                        const double u = draw(c.u_adm, c.lo + a, c.seed, day, STREAM_ADMISSION, (uint32_t)(c.lo + a), 0);
                        if (__ldg(c.p_adm + a) > u) { x = ST_H; ta[j] = day; w_ta |= 1u << j; atomicAdd(&s_ev[EV_ADM], 1u); }  // :213-216
                    }
                } else if (st == ST_H) {
                    if (ta[j] >= c.death_min) {                                   // :219
                        const double u = draw(c.u_death, c.lo + a, c.seed, day, STREAM_DEATH, (uint32_t)(c.lo + a), 0);
                        if (__ldg(c.p_death + a) > u) { x = ST_X; ta[j] = 0; w_ta |= 1u << j; atomicAdd(&s_ev[EV_DEATH], 1u); }  // :225-227, :259
                    }
                } else if (sus && srcv[j] >= 0) {                                 // :241-248: k_push found an infector
                    x = (st == ST_S) ? ST_I : ST_QI;
                    ti[j] = day; w_ti |= 1u << j;
                    ev_mask |= 1u << j;
                }
                // ---- E: the ordered rules of contagious3.py:250-259 ----
                const int n_inf = day - ti[j], n_q = day - tq[j];
                if (x == ST_QH && n_q == c.quarantine) x = ST_S;                                        // :250
                if (x == ST_QI && n_q == c.quarantine) x = ST_I;                                        // :252
                if (x == ST_I && n_inf == c.diagnosis) { tq[j] = day; w_tq |= 1u << j; }              // :253
                if ((x == ST_I || x == ST_QI) && n_inf == c.diagnosis) x = ST_D;                        // :254
                if (x == ST_D && n_inf == c.recover) x = ST_R;                                          // :255
                if (x == ST_H && n_inf == c.hospital_recover) { x = ST_R; ta[j] = 0; w_ta |= 1u << j; } // :256, :259
                if (x == ST_D && n_inf == c.diagnosis) atomicAdd(&s_ev[EV_NEW_DIAG], 1u);               // :261
                // ---- F (contagious3.py:261-270): the households diagnosed today were flagged yesterday ----
                if (any_nd && (x == ST_S || x == ST_I)) {
                    bool hit = hflag[j] & 1u;                                                           // :263
                    if (x == ST_I && !hit) hit = hflag[j] >> 1;                                         // :267 whole-row isin
                    if (hit && c.compliance < 1.0)
                        hit = keyed_uniform(c.seed, day, STREAM_COMPLIANCE, (uint32_t)(c.lo + a), 0) < c.compliance;
                    if (hit) { x = (x == ST_S) ? ST_QH : ST_QI; tq[j] = day; w_tq |= 1u << j; }         // :265-270
                }
                s[j] = x;
                if (x != st) { w_st = true; atomicAdd(&s_ev[EV_CHANGED], 1u); }
                if (in_set(M_ISO, x) && tq[j] == day) atomicAdd(&s_ev[EV_DAILY_QUAR], 1u);
                // tomorrow's phase A: snapshot byte, and the households that will see a diagnosis tomorrow
                ibw |= (uint32_t)infectious_byte(x, day + 1, ti[j]) << (8 * j);
                if (diagnosed_on(c, nullptr, a, x, ti[j], day + 1))
                    raise_household(c, a, c.hh_flag_next, c.hh_qflag_next, c.df_next);
            }
            // ---- G: integer statistics of the end-of-day state (contagious3.py:273-366) ----
            hist_add(h_end, x);
            if (in_set(M_EVER, x)) {
                const int k = day - ti[j];
                if (k >= 0 && k < HIST_LEN) {
                    const int64_t a = a0 + j;
                    if (k == 0) atomicAdd(&s_ev[EV_DAILY_INF], 1u);
                    atomicAdd(&s_hist[k], 1u);
                    if (c.sa_hist) { const int sa = c.sa[a]; if (sa >= 0) atomicAdd(&c.sa_hist[(size_t)sa * HIST_LEN + k], 1); }
                    if (c.io_mat && ((ev_mask >> j) & 1u)) {
                        // contagious3.py:356-366: row = area of today's infected, column = area of the infector
                        const int sa_u = c.sa[a], sa_i = c.sa[srcv[j] - c.lo];
                        if (sa_u >= 0 && sa_i >= 0) atomicAdd(&c.io_mat[(size_t)sa_u * c.S + sa_i], 1);
                    }
                }
                if (c.bcount && in_set(M_INF, x)) atomicAdd(&c.bcount[c.home[a0 + j]], 1);
            }
        }
        *reinterpret_cast<uint32_t*>(c.ib + c.lo + a0) = ibw;
        if (w_st) *reinterpret_cast<uchar4*>(c.status + a0) = make_uchar4(s[0], s[1], s[2], s[3]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {              // rare, and only the agents that changed (their neighbours' columns may not be loaded)
            if ((w_ti >> j) & 1u) c.t_inf[a0 + j] = (int16_t)ti[j];
            if ((w_tq >> j) & 1u) c.t_q[a0 + j] = (int16_t)tq[j];
            if ((w_ta >> j) & 1u) c.t_adm[a0 + j] = (int16_t)ta[j];
        }
    }
    // today's infection events: one slot claim per warp that has any (order is arbitrary; hosts sort)
    const unsigned any_ev = __ballot_sync(0xffffffffu, ev_mask != 0);
    if (any_ev) {
        const int mine = __popc(ev_mask);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) atomicAdd(&s_ev[EV_EVENTS], (unsigned)incl);
        if (c.ev) {
            int base = 0;
            if (lane == 31) base = atomicAdd(&c.df[DF_N_EVENTS], incl);
            base = __shfl_sync(0xffffffffu, base, 31) + incl - mine;
            const int srcv[4] = {src4.x, src4.y, src4.z, src4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if ((ev_mask >> j) & 1u) {
                    c.ev[base++] = make_int2((int)(c.lo + a0 + j), srcv[j]);
                }
        }
    }
    if (VARIANT == 2) return;
    // ---- block reduction: packed histograms by REDUX, then one global atomic per non-zero counter ----
    {
        const uint32_t e0 = __reduce_add_sync(0xffffffffu, h_end[0]), e1 = __reduce_add_sync(0xffffffffu, h_end[1]);
        const uint32_t b0 = __reduce_add_sync(0xffffffffu, h_start);
        if (lane < 10) {
            const uint32_t w = lane < 4 ? e0 : (lane < 8 ? e1 : b0);
            const uint32_t v = (w >> (8 * (lane & 3))) & 0xFFu;
            if (v) atomicAdd(&s_cnt[lane], v);
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        // warp 0 publishes everything, fences, and takes the ticket
        const int slot = blockIdx.x & (N_SLOTS - 1);
        if (lane < 10 && s_cnt[lane])
            atomicAdd(&c.part[(size_t)(RAW_END + lane) * N_SLOTS + slot], (unsigned long long)s_cnt[lane]);
        if (lane < EV_LEN && s_ev[lane]) {
            const int idx[EV_LEN] = {ST_DAILY_HOSP, ST_DAILY_DEATHS, ST_DAILY_INF, ST_DAILY_QUAR, ST_NEW_DIAG, ST_N_EVENTS, ST_CHANGED};
            atomicAdd(&c.part[(size_t)idx[lane] * N_SLOTS + slot], (unsigned long long)s_ev[lane]);
        }
        if (lane < HIST_LEN && s_hist[lane])
            atomicAdd(&c.part[(size_t)(HIST0 + lane) * N_SLOTS + slot], (unsigned long long)s_hist[lane]);
        if (VARIANT == 0) {
            __threadfence();
            __syncwarp();
            if (lane == 0) s_last = (atomicAdd(c.ticket, 1u) == gridDim.x - 1);
        }
    }
    if (VARIANT != 0) return;
    __syncthreads();
    if (s_last) {
        // the last block folds the spread partial sums: part[counter][slot], one warp per counter, two coalesced loads
        __threadfence();
        unsigned long long t[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            unsigned long long* p = c.part + (size_t)((threadIdx.x >> 5) + 8 * q) * N_SLOTS;
            t[q] = __ldcg(p + lane) + __ldcg(p + lane + 32);
            p[lane] = 0ull;
            p[lane + 32] = 0ull;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
#pragma unroll
            for (int o = 16; o; o >>= 1) t[q] += __shfl_xor_sync(0xffffffffu, t[q], o);
            if (lane == 0) s_fold[(threadIdx.x >> 5) + 8 * q] = (long long)t[q];
        }
        __syncthreads();
        if (threadIdx.x < N_STATS) {
            const long long* r = s_fold + RAW_END;        // QI S I QH D H R X
            long long v = s_fold[threadIdx.x];
            switch (threadIdx.x) {
                case ST_ACTIVE: v = r[2] + r[0] + r[4] + r[5]; break;
                case ST_RECOVERED: v = r[6]; break;
                case ST_QUARANTINED: v = r[3] + r[0] + r[4]; break;
                case ST_HOSPITALIZED: v = r[5]; break;
                case ST_DEAD: v = r[7]; break;
                case ST_EVER_INFECTED: v = r[2] + r[0] + r[4] + r[5] + r[6] + r[7]; break;
                case ST_SUS0: v = s_fold[RAW_SUS0]; break;
                case ST_INF0: v = s_fold[RAW_INF0]; break;
                case ST_COUNT0 + 0: v = r[1]; break;
                case ST_COUNT0 + 1: v = r[2]; break;
                case ST_COUNT0 + 2: v = r[3]; break;
                case ST_COUNT0 + 3: v = r[0]; break;
                case ST_COUNT0 + 4: v = r[4]; break;
                case ST_COUNT0 + 5: v = r[5]; break;
                case ST_COUNT0 + 6: v = r[6]; break;
                case ST_COUNT0 + 7: v = r[7]; break;
                default: break;
            }
            if (threadIdx.x >= RAW_END) v = 0;
            c.stats[threadIdx.x] = v;
        }
        if (threadIdx.x == 0) *c.ticket = 0u;
    }
}

__global__ void __launch_bounds__(256) k_agent(const DevCtx c, const int day) { agent_body<0>(c, day); }

// phase A alone (contagious3.py:181-202): the day-start snapshot byte of every owned agent and the households that see
// a diagnosis on `day`.  The host clears hh_flag / hh_qflag / df first.
__global__ void __launch_bounds__(256) k_snapshot(const DevCtx c, const int day, uint8_t* hh_flag, uint8_t* hh_qflag, int32_t* df)
{
    const int64_t a0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    if (a0 >= c.n_pad) return;
    const uchar4 sv = *reinterpret_cast<const uchar4*>(c.status + a0);
    const short4 ti4 = *reinterpret_cast<const short4*>(c.t_inf + a0);
    const uint32_t st[4] = {sv.x, sv.y, sv.z, sv.w};
    const int ti[4] = {ti4.x, ti4.y, ti4.z, ti4.w};
    uint8_t ibn[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        ibn[j] = infectious_byte(st[j], day, ti[j]);
        if (diagnosed_on(c, c.u_adm, a0 + j, st[j], ti[j], day)) raise_household(c, a0 + j, hh_flag, hh_qflag, df);
    }
    *reinterpret_cast<uchar4*>(c.ib + c.lo + a0) = make_uchar4(ibn[0], ibn[1], ibn[2], ibn[3]);
}

}  // namespace dys
