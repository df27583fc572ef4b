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
// k_push.  Persistent, warp-autonomous.  A warp scans the snapshot bytes of 128 halo agents per step (one 32-bit load
// per lane, the next step's load already in flight), compacts the infectious ones into its own shared-memory frontier
// queue, and whenever 32 rows are queued expands them together: the 32 transposed rows are flattened (warp prefix sum
// of their lengths) so every lane works on consecutive entries however the infectious agents are spread; four entries
// per lane are in flight (one coalesced 4-byte read + one 1-byte status gather each).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

constexpr int PUSH_WARPS = 8;
constexpr int PUSH_CHUNK = 128;                    // halo agents scanned per warp step (4 snapshot bytes per lane): small
                                                   // steps spread the infectious rows over all resident warps
constexpr int PUSH_QCAP = PUSH_CHUNK + 32;
constexpr int PUSH_PCAP = 256;

struct PushWarpSmem {
    uint32_t q_idx[PUSH_QCAP];                     // halo-relative index of a queued infectious agent
    uint8_t q_ib[PUSH_QCAP];                       // its snapshot byte
    int off[36];                                   // exclusive prefix of the 32 expanded rows' lengths
    int64_t t0[32];
    unsigned long long pairs[PUSH_PCAP];           // exposed pairs awaiting their Bernoulli: entry | row << 48 | iso_u << 56
};

__device__ __forceinline__ void push_body(const DevCtx& c, const int day)
{
    __shared__ double s_pdf[64];
    __shared__ unsigned int s_acc[8];
    __shared__ PushWarpSmem s_w[PUSH_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    PushWarpSmem& w = s_w[wid];

    // scratch that k_agent fills today, and the flags it raises for tomorrow
    grid_clear(c.sa_hist, c.S_out * HIST_LEN);
    grid_clear(c.bcount, c.B_out);
    grid_clear(c.io_mat, c.io_mat ? c.S * c.S : 0);
    grid_clear16(c.hh_flag_next, c.H);
    grid_clear16(c.hh_qflag_next, c.H);
    if (blockIdx.x == 0 && tid < DF_LEN) c.df_next[tid] = 0;

    if (tid < 64) s_pdf[tid] = c.pdf[tid];
    if (tid < 8) s_acc[tid] = 0;
    __syncthreads();
    const bool any_closed = c.n_closed != 0;
    unsigned int cnt[4] = {0, 0, 0, 0};   // entries scanned, candidates, exposed, hits

    // settle the first n queued pairs (susceptible u, infectious i) that share a building: probability, Bernoulli, mailbox.
    // Dense: one pair per lane, so the two 8-byte gathers and the Philox rounds are issued once per 32 pairs.
    auto settle = [&](const int q0, const int n) {
        for (int k0 = 0; k0 < n; k0 += 32) {
            const int k = k0 + lane;
            if (k >= n) continue;
            const unsigned long long rec = w.pairs[k];
            const int64_t e = (int64_t)(rec & 0xFFFFFFFFFFFFull);
            const int r = (int)((rec >> 48) & 0x3Fu);
            const bool iso_u = (rec >> 56) & 1u;
            const uint32_t cw = __ldg(c.tcolx + e), ib_i = w.q_ib[q0 + r];
            const uint32_t ul = cw & COL_MASK;
            const int64_t ug = c.lo + ul, ig = c.rt_lo + w.q_idx[q0 + r];
            if (any_closed && !exposed_slow(c.routine, c.open, c.V, ug - c.rt_lo, ig - c.rt_lo, iso_u, (ib_i & IB_ISO) != 0)) continue;
            ++cnt[2];
            // contagious3.py:232-238: ((ip[u,i] * risk[u]) * pdf(days since i's infection)) * norm_factor, each product
            // rounded separately (no add follows, so no FMA can form)
            double p = __dmul_rn(__ldg(c.tval + e), __ldg(c.risk + ul));
            p = __dmul_rn(p, s_pdf[ib_i & IB_DAYS]);
            p = __dmul_rn(p, c.norm_factor);
            if (c.pair_prob) c.pair_prob[(size_t)ug * c.n_glob + ig] = p;
            const double u = draw(c.u_pair, ug * c.n_glob + ig, c.seed, day, STREAM_INFECTION, (uint32_t)ug, (uint32_t)ig);
            if (p > u) { ++cnt[3]; atomicMax(&c.src[ul], (int)ig); }               // :240, :248
        }
        __syncwarp();
    };

    // expand nb (<= 32) queued rows starting at queue position q0
    auto expand = [&](const int q0, const int nb) {
        int64_t t0 = 0;
        int len = 0;
        if (lane < nb) { const uint32_t il = w.q_idx[q0 + lane]; t0 = __ldg(c.tptr + il); len = (int)(__ldg(c.tptr + il + 1) - t0); }
        int incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        w.off[lane] = incl - len;
        if (lane == 31) w.off[32] = total;
        w.t0[lane] = t0;
        __syncwarp();
        cnt[0] += len;
        int np = 0;                                        // queued pairs (warp-uniform)
        for (int f0 = 0; f0 < total; f0 += 128) {
            int64_t e[4];
            int r[4];
            uint32_t cw[4], st_u[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int f = f0 + 32 * k + lane;
                int lo = 0, hi = 32;                       // off[lo] <= f < off[hi]; the last such lo is the row
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                    const int mid = (lo + hi) >> 1;
                    if (w.off[mid] <= f) lo = mid; else hi = mid;
                }
                r[k] = lo;
                e[k] = w.t0[lo] + (f - w.off[lo]);
                cw[k] = f < total ? __ldg(c.tcolx + e[k]) : 0u;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                st_u[k] = (f0 + 32 * k + lane) < total ? (uint32_t)c.status[cw[k] & COL_MASK] : (uint32_t)ST_PAD;   // day-start status
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                // cheap part of the exposure test (static class x today's isolation); survivors are queued
                const bool sus = in_set(M_SUS, st_u[k]);
                cnt[1] += sus;
                const uint32_t cls = cw[k] >> CLS_SHIFT;
                const bool iso_u = (st_u[k] == ST_QH), iso_i = w.q_ib[q0 + r[k]] & IB_ISO;
                const bool ex = sus && ((cls & CLS_HH) | ((cls & CLS_HO) && !iso_i) | ((cls & CLS_OH) && !iso_u) |
                                        ((cls & CLS_OO) && !iso_u && !iso_i));
                const unsigned m = __ballot_sync(0xffffffffu, ex);
                if (ex) {
                    w.pairs[np + __popc(m & lt_mask)] = (unsigned long long)e[k] | ((unsigned long long)r[k] << 48) |
                                                        ((unsigned long long)iso_u << 56);
                    prefetch_l2(c.tval + e[k]);                 // both are read when the pair is settled
                    prefetch_l2(c.risk + (cw[k] & COL_MASK));
                }
                np += __popc(m);
            }
            __syncwarp();
            if (np > PUSH_PCAP - 128) { settle(q0, np); np = 0; }
        }
        if (np > 0) settle(q0, np);
        __syncwarp();
    };

    const int64_t n_chunks = (c.rt_n + PUSH_CHUNK - 1) / PUSH_CHUNK;
    const int64_t stride = (int64_t)gridDim.x * PUSH_WARPS;
    int nq = 0;                                            // queued rows (warp-uniform)
    const uint32_t* ib4 = reinterpret_cast<const uint32_t*>(c.ib_rd + c.rt_lo);  // rt_lo is a multiple of 1024
    // scan the snapshot bytes of halo agents [ch0, ch1) * 128, queueing the infectious ones
    auto scan = [&](const int64_t ch0, const int64_t ch1) {
        int64_t ch = ch0 + (int64_t)blockIdx.x * PUSH_WARPS + wid;
        uint32_t v = 0;
        if (ch < ch1) v = ib4[ch * (PUSH_CHUNK / 4) + lane];
        for (; ch < ch1; ch += stride) {
            uint32_t vn = 0;
            if (ch + stride < ch1) vn = ib4[(ch + stride) * (PUSH_CHUNK / 4) + lane];   // next step's bytes, in flight
            const int mine = __popc(v & 0x80808080u);
            if (__ballot_sync(0xffffffffu, mine != 0)) {
                int incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                int pos = nq + incl - mine;
                const uint32_t il0 = (uint32_t)(ch * PUSH_CHUNK + lane * 4);
                if (mine) {
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const uint32_t byte = (v >> (8 * b)) & 0xFFu;
                        if ((byte & IB_INF) && il0 + b < (uint32_t)c.rt_n) {
                            w.q_idx[pos] = il0 + b; w.q_ib[pos] = (uint8_t)byte; ++pos;
                            prefetch_l2(c.tptr + il0 + b);      // the row extent is needed when the queue is expanded
                        }
                    }
                }
                nq += __shfl_sync(0xffffffffu, incl, 31);
                __syncwarp();
                while (nq >= 32) {     // expand from the back of the queue: nothing has to move
                    nq -= 32;
                    expand(nq, 32);
                }
            }
            v = vn;
        }
    };
    if (c.n_peers > 1) {
        // peer mode: the owned agents' bytes are local and final, so their rows are pushed first; the halo agents'
        // bytes were stored into this GPU's buffer by their owners, who then raised flag[r] = day -- the exchange
        // latency hides behind the local part of the push.  (Persistent grid: all CTAs resident, spinning is safe.)
        const int64_t own0 = (c.lo - c.rt_lo) / PUSH_CHUNK, own1 = (c.lo - c.rt_lo + c.n_loc + PUSH_CHUNK - 1) / PUSH_CHUNK;
        scan(own0, own1);
        if (tid < c.n_peers && ((c.wait_mask >> tid) & 1u)) {      // only the ranks whose agents are in this rank's halo
            const volatile int32_t* f = c.flag_wait + tid;
            while (*f < day) __nanosleep(32);
            __threadfence_system();
        }
        __syncthreads();
        scan(0, own0);
        scan(own1, n_chunks);
    } else {
        scan(0, n_chunks);
    }
    if (nq > 0) expand(0, nq);
    const int idx[4] = {ST_EDGES_SCANNED, ST_CANDIDATES, ST_EXPOSED, ST_HITS};
    __syncthreads();
    block_accumulate<4>(c, idx, cnt, s_acc);
}

__global__ void __launch_bounds__(PUSH_WARPS * 32) k_push(const DevCtx c, const int day) { push_body(c, day); }

// ---------------------------------------------------------------------------------------------------------------
// k_agent.  A block owns 1024 consecutive agents, in two phases.
//   sweep  : 4 agents per thread, vector loads of the status word, the infector mailbox and (on days with a
//            diagnosis) the household flags.  An agent to whom nothing can happen today -- susceptible, nobody infected
//            it, household not flagged; or recovered -- only feeds the packed status histogram.  Every other agent's
//            local index is compacted (ballot + prefix popcount) into a shared-memory work list.
//   settle : the work list is processed densely, one agent per thread: B, C, applying D, the ordered rules of E, F, G,
//            tomorrow's snapshot byte and household flags.  Divergent per-agent logic therefore costs instructions in
//            proportion to the agents that need it, not to the population.
// Compartment counts are reduced as packed 8-bit histograms (REDUX); rare events go to shared-memory counters; the last
// block folds the spread partial sums into the day's stats block.
// ---------------------------------------------------------------------------------------------------------------
// packed 8-bit histogram field of a status code: QI S I QH | D H R X   (4 agents x 32 lanes = 128 < 256 per field)
__device__ __forceinline__ uint32_t hist_idx(uint32_t st) { return (st == ST_QI) ? 0u : (st >> 1); }
constexpr uint32_t M_F = M_SUS | sbit(ST_I) | sbit(ST_QI);      // can be in status 1 or 2 when F runs
enum { EV_ADM = 0, EV_DEATH, EV_DAILY_INF, EV_DAILY_QUAR, EV_NEW_DIAG, EV_EVENTS, EV_CHANGED, EV_LEN };
// internal partial-sum rows (folded into the public stats by the last block)
constexpr int RAW_END = 53;      // [53..60] end-of-day status counts in hist_idx order
constexpr int RAW_SUS0 = 61, RAW_INF0 = 62;
constexpr int ST_COUNT0 = 20;    // public: stats[20..27] = end-of-day agents in S, I, QH, QI, D, H, R, X
constexpr int AGENT_TILE = 1024;

// one non-idle agent, start to end of day (contagious3.py:205-270 for this agent, then its share of :273-366)
__device__ __forceinline__ void settle_agent(const DevCtx& c, const int day, const int64_t a, const bool any_nd,
                                             const uint32_t st, const int mailbox, unsigned int* s_cnt, unsigned int* s_ev,
                                             unsigned int* s_hist, bool& infected_today, int& infector)
{
    // every column this agent can need is requested before any of it is used: the per-agent latency is two dependent
    // round trips (columns, then the household flags) instead of one per rule
    int ti = c.t_inf[a];
    int tq = c.t_q[a];
    const int h = any_nd ? __ldg(c.hh + a) : 0;
    const int sa_a = c.sa_hist ? __ldg(c.sa + a) : -1;
    const int home_a = c.bcount ? __ldg(c.home + a) : 0;
    bool f_hh = false, f_q = false;
    if (any_nd) {
        f_hh = __ldg(c.hh_flag + h) != 0;
        f_q = (__ldg(c.hh_qflag + h) != 0) || (c.hh_always && __ldg(c.hh_always + h));
    }
    uint32_t x = st;
    if (!in_set(M_EVER, st)) ti = 0;
    bool w_ti = false, w_tq = false;
    infected_today = false;
    // ---- B, C (contagious3.py:205-227) on the day-start status; D's result from the mailbox ----
    if (in_set(M_INF, st)) {
        atomicAdd(&s_cnt[9], 1u);
        if (ti >= c.adm_min && ti <= c.adm_max) {                            // :205 window on the ABSOLUTE infection day
            const double u = draw(c.u_adm, c.lo + a, c.seed, day, STREAM_ADMISSION, (uint32_t)(c.lo + a), 0);
            if (__ldg(c.p_adm + a) > u) { x = ST_H; c.t_adm[a] = (int16_t)day; atomicAdd(&s_ev[EV_ADM], 1u); }   // :213-216
        }
    } else if (st == ST_H) {
        if (c.t_adm[a] >= c.death_min) {                                     // :219
            const double u = draw(c.u_death, c.lo + a, c.seed, day, STREAM_DEATH, (uint32_t)(c.lo + a), 0);
            if (__ldg(c.p_death + a) > u) { x = ST_X; c.t_adm[a] = 0; atomicAdd(&s_ev[EV_DEATH], 1u); }          // :225-227, :259
        }
    } else if (in_set(M_SUS, st)) {
        atomicAdd(&s_cnt[8], 1u);
        infector = mailbox;
        if (infector >= 0) {                                                 // :241-248: k_push found an infector
            x = (st == ST_S) ? ST_I : ST_QI;
            ti = day; w_ti = true;
            infected_today = true;
        }
    }
    // ---- E: the ordered rules of contagious3.py:250-259 ----
    const int n_inf = day - ti, n_q = day - tq;
    if (x == ST_QH && n_q == c.quarantine) x = ST_S;                                        // :250
    if (x == ST_QI && n_q == c.quarantine) x = ST_I;                                        // :252
    if (x == ST_I && n_inf == c.diagnosis) { tq = day; w_tq = true; }                       // :253
    if ((x == ST_I || x == ST_QI) && n_inf == c.diagnosis) x = ST_D;                        // :254
    if (x == ST_D && n_inf == c.recover) x = ST_R;                                          // :255
    if (x == ST_H && n_inf == c.hospital_recover) { x = ST_R; c.t_adm[a] = 0; }             // :256, :259
    if (x == ST_D && n_inf == c.diagnosis) atomicAdd(&s_ev[EV_NEW_DIAG], 1u);               // :261
    // ---- F (contagious3.py:261-270): the households diagnosed today were flagged yesterday ----
    if (any_nd && (x == ST_S || x == ST_I)) {
        bool hit = f_hh;                                                                    // :263
        if (x == ST_I && !hit) hit = f_q;                                                   // :267 whole-row isin
        if (hit && c.compliance < 1.0)
            hit = keyed_uniform(c.seed, day, STREAM_COMPLIANCE, (uint32_t)(c.lo + a), 0) < c.compliance;
        if (hit) { x = (x == ST_S) ? ST_QH : ST_QI; tq = day; w_tq = true; }                // :265-270
    }
    if (w_ti) c.t_inf[a] = (int16_t)ti;
    if (w_tq) c.t_q[a] = (int16_t)tq;
    if (x != st) { c.status[a] = (uint8_t)x; atomicAdd(&s_ev[EV_CHANGED], 1u); }
    if (in_set(M_ISO, x) && tq == day) atomicAdd(&s_ev[EV_DAILY_QUAR], 1u);
    // ---- G: this agent's share of the integer statistics (contagious3.py:273-366) ----
    atomicAdd(&s_cnt[hist_idx(x)], 1u);
    if (in_set(M_EVER, x)) {
        const int k = day - ti;
        if (k >= 0 && k < HIST_LEN) {
            if (k == 0) atomicAdd(&s_ev[EV_DAILY_INF], 1u);
            atomicAdd(&s_hist[k], 1u);
            if (c.sa_hist && sa_a >= 0) atomicAdd(&c.sa_hist[(size_t)sa_a * HIST_LEN + k], 1);
            if (c.io_mat && infected_today) {
                // contagious3.py:356-366: row = area of today's infected, column = area of the infector
                const int sa_u = __ldg(c.sa + a), sa_i = __ldg(c.sa + (infector - c.lo));
                if (sa_u >= 0 && sa_i >= 0) atomicAdd(&c.io_mat[(size_t)sa_u * c.S + sa_i], 1);
            }
        }
        if (c.bcount && in_set(M_INF, x)) atomicAdd(&c.bcount[home_a], 1);
    }
    // ---- tomorrow's phase A: snapshot byte, and the households that will see a diagnosis tomorrow ----
    const uint8_t ibn = infectious_byte(x, day + 1, ti);
    c.ib_wr[0][c.lo + a] = ibn;
    if (diagnosed_on(c, nullptr, a, x, ti, day + 1)) raise_household(c, a, c.hh_flag_next, c.hh_qflag_next, c.df_next);
}

// One block folds the spread partial sums (part[counter][slot]: one warp per counter, two coalesced loads, shuffle tree),
// derives the compartment statistics from the status counts and re-zeroes the partials for the next day.
__device__ __forceinline__ void fold_stats(const DevCtx& c)
{
    __shared__ long long s_fold[N_STATS];
    const int tid = threadIdx.x, lane = tid & 31;
    __threadfence();
    unsigned long long t[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        unsigned long long* p = c.part + (size_t)((tid >> 5) + 8 * q) * N_SLOTS;
        t[q] = __ldcg(p + lane) + __ldcg(p + lane + 32);
        p[lane] = 0ull;
        p[lane + 32] = 0ull;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
#pragma unroll
        for (int o = 16; o; o >>= 1) t[q] += __shfl_xor_sync(0xffffffffu, t[q], o);
        if (lane == 0) s_fold[(tid >> 5) + 8 * q] = (long long)t[q];
    }
    __syncthreads();
    if (tid < N_STATS) {
        const long long* r = s_fold + RAW_END;        // QI S I QH D H R X
        long long v = s_fold[tid];
        switch (tid) {
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
        if (tid >= RAW_END) v = 0;
        c.stats[tid] = v;
    }
}

template <int VARIANT>   // 0 = k_agent (last block folds); 3 = k_days (block 0 folds after the grid barrier);
                         // 1, 2 = tools/kbench.cu ablations (no fold / no reduction at all)
__device__ __forceinline__ void agent_body(const DevCtx& c, const int day)
{
    __shared__ unsigned int s_hist[HIST_LEN];
    __shared__ unsigned int s_ev[EV_LEN];
    __shared__ unsigned int s_cnt[10];            // 8 end-of-day status counts (hist_idx order), sus0, inf0
    __shared__ unsigned int s_nlist;
    __shared__ unsigned short s_list[AGENT_TILE];
    __shared__ uint32_t s_stw[AGENT_TILE / 4];      // the tile's status words, for the settle phase
    __shared__ __align__(16) int s_src[AGENT_TILE]; // the tile's mailbox (infector) column (written 16 bytes at a time)
    __shared__ bool s_last;
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    if (tid < HIST_LEN) s_hist[tid] = 0;
    if (tid < EV_LEN) s_ev[tid] = 0;
    if (tid < 10) s_cnt[tid] = 0;
    __syncthreads();
    const bool any_nd = c.df[DF_ANY_ND] != 0;
    const bool r_idle = c.recover >= HIST_LEN && c.hospital_recover >= HIST_LEN;   // a recovered agent left the histogram
    // one tile per block when launched as k_agent (grid = tiles); a persistent grid (k_days) strides over the tiles
    for (int64_t tile = blockIdx.x; tile * AGENT_TILE < c.n_pad; tile += gridDim.x) {
    const int64_t tile0 = tile * AGENT_TILE;
    const int64_t a0 = tile0 + tid * 4;
    if (tid == 0) s_nlist = 0;
    __syncthreads();

    // ---- sweep ----
    {
        // all three columns are requested at once (the mailbox and household index of a susceptible agent are needed
        // on almost every day of an epidemic), then the household flags
        const uint32_t sw = *reinterpret_cast<const uint32_t*>(c.status + a0);
        const int4 src4 = *reinterpret_cast<const int4*>(c.src + a0);        // the mailbox k_push wrote into
        int4 hh4 = make_int4(0, 0, 0, 0);
        if (any_nd) hh4 = __ldg(reinterpret_cast<const int4*>(c.hh + a0));
        const uint32_t st0[4] = {sw & 0xFFu, (sw >> 8) & 0xFFu, (sw >> 16) & 0xFFu, sw >> 24};
        s_stw[tid] = sw;
        *reinterpret_cast<int4*>(&s_src[tid * 4]) = src4;
        uint32_t m_f = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) m_f |= (uint32_t)(st0[j] == ST_S) << j;
        const int srcv[4] = {src4.x, src4.y, src4.z, src4.w};
        uint32_t flagged = 0;                      // susceptible agents whose household sees a diagnosis today
        if (any_nd && m_f) {
            const int hv[4] = {hh4.x, hh4.y, hh4.z, hh4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if ((m_f >> j) & 1u) flagged |= (uint32_t)(__ldg(c.hh_flag + hv[j]) != 0) << j;
        }
        uint32_t work = 0, n_idle_s = 0, n_idle_r = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t st = st0[j];
            const bool idle_s = (st == ST_S) && srcv[j] < 0 && !((flagged >> j) & 1u);
            const bool idle_r = (st == ST_R) && r_idle;
            n_idle_s += idle_s;
            n_idle_r += idle_r;
            work |= (uint32_t)(!idle_s && !idle_r && st != ST_PAD) << j;
        }
        // idle agents: end-of-day status == day-start status; their snapshot byte is 0
        if (VARIANT != 2) {
            const uint32_t packed = n_idle_s | (n_idle_r << 8);
            const uint32_t tot = __reduce_add_sync(0xffffffffu, packed);
            if (lane == 0 && tot) {
                if (tot & 0xFFu) { atomicAdd(&s_cnt[1], tot & 0xFFu); atomicAdd(&s_cnt[8], tot & 0xFFu); }
                if (tot >> 8) atomicAdd(&s_cnt[6], tot >> 8);
            }
        }
        // tomorrow's snapshot byte of an idle agent is 0; the settle phase overwrites the bytes of the agents it handles
        *reinterpret_cast<uint32_t*>(c.ib_wr[0] + c.lo + a0) = 0u;
        // compact the agents that need the settle phase
        const int mine = __popc(work);
        const unsigned any = __ballot_sync(0xffffffffu, mine != 0);
        if (any) {
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            int base = 0;
            if (lane == 31) base = (int)atomicAdd(&s_nlist, (unsigned)incl);
            base = __shfl_sync(0xffffffffu, base, 31) + incl - mine;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if ((work >> j) & 1u) s_list[base++] = (unsigned short)(tid * 4 + j);
        }
    }
    __syncthreads();

    // ---- settle ----
    const int n_list = (int)s_nlist;
    for (int k0 = 0; k0 < n_list; k0 += 256) {
        const int k = k0 + tid;
        const bool valid = k < n_list;
        bool infected_today = false;
        int infector = -1;
        int64_t a = 0;
        if (valid) {
            const int li = s_list[k];
            a = tile0 + li;
            settle_agent(c, day, a, any_nd, (s_stw[li >> 2] >> (8 * (li & 3))) & 0xFFu, s_src[li], s_cnt, s_ev, s_hist,
                         infected_today, infector);
        }
        // today's infection events: one slot claim per warp that has any (order is arbitrary; hosts sort)
        const unsigned evm = __ballot_sync(0xffffffffu, infected_today);
        if (evm) {
            if (lane == 0) atomicAdd(&s_ev[EV_EVENTS], (unsigned)__popc(evm));
            if (c.ev) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&c.df[DF_N_EVENTS], __popc(evm));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (infected_today) c.ev[base + __popc(evm & lt_mask)] = make_int2((int)(c.lo + a), infector);
            }
        }
    }
    __syncthreads();      // s_list / s_nlist are free for the next tile
    }
    if (VARIANT == 2) return;
    __syncthreads();
    if (tid < 32) {
        // warp 0 publishes everything, fences, and takes the ticket
        const int slot = blockIdx.x & (N_SLOTS - 1);
        if (lane < 10 && s_cnt[lane])
            atomicAdd(&c.part[(size_t)(RAW_END + lane) * N_SLOTS + slot], (unsigned long long)s_cnt[lane]);
        if (lane < EV_LEN && s_ev[lane]) {
            const int row = lane == EV_ADM ? ST_DAILY_HOSP : lane == EV_DEATH ? ST_DAILY_DEATHS : lane == EV_DAILY_INF ? ST_DAILY_INF
                          : lane == EV_DAILY_QUAR ? ST_DAILY_QUAR : lane == EV_NEW_DIAG ? ST_NEW_DIAG : lane == EV_EVENTS ? ST_N_EVENTS
                          : ST_CHANGED;
            atomicAdd(&c.part[(size_t)row * N_SLOTS + slot], (unsigned long long)s_ev[lane]);
        }
        if (lane < HIST_LEN && s_hist[lane])
            atomicAdd(&c.part[(size_t)(HIST0 + lane) * N_SLOTS + slot], (unsigned long long)s_hist[lane]);
        if (VARIANT == 0) {
            __threadfence();
            __syncwarp();
            if (lane == 0) s_last = (atomicAdd(c.ticket, 1u) == gridDim.x - 1);      // one ticket per block, after all its tiles
        }
    }
    if (VARIANT != 0) return;
    __syncthreads();
    if (s_last) {
        fold_stats(c);      // the last block to finish
        if (tid == 0) *c.ticket = 0u;
    }
}

__global__ void __launch_bounds__(256) k_agent(const DevCtx c, const int day) { agent_body<0>(c, day); }

// ---------------------------------------------------------------------------------------------------------------
// k_days: several consecutive days in ONE cooperative launch (persistent grid, all CTAs resident).  Per day: push_body,
// grid barrier, agent_body, grid barrier -- the same device code as k_push / k_agent, minus two launches a day.  At a
// million agents a day is ~20 MB of traffic, i.e. launch latency is the larger part of a day done as separate kernels.
// The per-day buffers (output block, household flags, flag block, snapshot parity, open flags) come from DayPlan.
// ---------------------------------------------------------------------------------------------------------------
// peer mode: the owned slice of freshly written snapshot bytes (ib_wr[0]) is stored into the same place of every peer
// GPU's buffer (ib_wr[1..]) with coalesced 16-byte P2P stores over NVLink; the flag is raised afterwards (k_signal, or
// the tail of a k_days day) once every block's stores are fenced
__device__ __forceinline__ void publish_body(const DevCtx& c)
{
    // only the part of the owned slice that lies inside the peer's halo travels (pub_lo / pub_n, 16-byte granules)
    for (int p = 1; p < c.n_wr; ++p) {
        const uint4* src = reinterpret_cast<const uint4*>(c.ib_wr[0] + c.pub_lo[p]);
        uint4* dst = reinterpret_cast<uint4*>(c.ib_wr[p] + c.pub_lo[p]);
        const int64_t n16 = c.pub_n[p] >> 4;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
    }
    __threadfence_system();
}

__global__ void __launch_bounds__(256) k_publish(const DevCtx c) { publish_body(c); }

constexpr int MAX_FUSED_DAYS = 14;
constexpr int PUBLISH_BLOCKS = 8;          // the halo part of a slice is small (one community per neighbour)
struct DayPlan {
    int32_t first_day, n_days;
    int32_t n_closed[MAX_FUSED_DAYS];
    const uint8_t* open[MAX_FUSED_DAYS];
    unsigned char* out[MAX_FUSED_DAYS];             // the day's output block
    size_t off_sa, off_bc, off_ev;                  // 0 = that part was not requested
    uint8_t* hh_flag[2];
    uint8_t* hh_qflag[2];
    int32_t* df[2];
    uint8_t* ib[2];                                 // own snapshot buffers by day parity
    uint8_t* peer_ib[8][2];                         // peer mode: every other rank's, by day parity
    int32_t first_parity;                           // step parity of first_day (household flags, flag block)
    int32_t want_sa, want_bc, want_ev;
    unsigned long long* trace;                      // developer tracing (tools/kbench.cu): globaltimer at phase boundaries, or null
};

__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& epoch)
{
    // all CTAs are resident (cooperative launch); one arrival per CTA, monotone counter, no reset needed
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        ++epoch;
        const unsigned int target = epoch * gridDim.x;
        while (*reinterpret_cast<volatile unsigned int*>(counter) < target) __nanosleep(32);
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256, 4) k_days(const DevCtx c0, const DayPlan plan, unsigned int* barrier_counter)
{
    __shared__ DevCtx s_c;               // the day's view of the context: read by both bodies like kernel parameters
    unsigned int epoch = 0;
    for (int k = 0; k < plan.n_days; ++k) {
        const int day = plan.first_day + k;
        if (threadIdx.x == 0) {
            DevCtx& c = s_c;
            c = c0;
            const int par = (plan.first_parity + k) & 1;
            c.open = plan.open[k];
            c.n_closed = plan.n_closed[k];
            c.stats = reinterpret_cast<int64_t*>(plan.out[k]);
            c.sa_hist = plan.want_sa ? reinterpret_cast<int32_t*>(plan.out[k] + plan.off_sa) : nullptr;
            c.bcount = plan.want_bc ? reinterpret_cast<int32_t*>(plan.out[k] + plan.off_bc) : nullptr;
            c.ev = plan.want_ev ? reinterpret_cast<int2*>(plan.out[k] + plan.off_ev) : nullptr;
            c.hh_flag = plan.hh_flag[par]; c.hh_qflag = plan.hh_qflag[par];
            c.hh_flag_next = plan.hh_flag[par ^ 1]; c.hh_qflag_next = plan.hh_qflag[par ^ 1];
            c.df = plan.df[par]; c.df_next = plan.df[par ^ 1];
            c.ib_rd = plan.ib[day & 1];
            const int wr = (day + 1) & 1;
            c.n_wr = 0;
            c.ib_wr[c.n_wr++] = plan.ib[wr];
            for (int r = 0; r < c.n_peers; ++r)
                if (c.n_peers > 1 && r != c.rank) c.ib_wr[c.n_wr++] = plan.peer_ib[r][wr];
        }
        __syncthreads();
        auto stamp = [&](int i) {
            if (plan.trace && blockIdx.x == 0 && threadIdx.x == 0) {
                unsigned long long t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                plan.trace[k * 8 + i] = t;
            }
        };
        stamp(0);
        push_body(s_c, day);
        stamp(1);
        grid_barrier(barrier_counter, epoch);
        stamp(2);
        agent_body<3>(s_c, day);
        stamp(3);
        grid_barrier(barrier_counter, epoch);
        stamp(4);
        if (blockIdx.x == gridDim.x - 1) fold_stats(s_c);      // every block's partial sums are in; the others move on
        if (c0.n_peers > 1 && blockIdx.x < PUBLISH_BLOCKS) {
            // a few CTAs ship tomorrow's bytes to the peers while the others already start the next day (whose push
            // waits for the peers' flags anyway); the last shipper raises this rank's flag on every GPU
            publish_body(s_c);
            __syncthreads();
            __shared__ bool s_pub_last;
            if (threadIdx.x == 0) s_pub_last = (atomicAdd(barrier_counter + 1, 1u) == PUBLISH_BLOCKS - 1);
            __syncthreads();
            if (s_pub_last) {
                if (threadIdx.x == 0) barrier_counter[1] = 0u;
                if (threadIdx.x < c0.n_peers && ((c0.signal_mask >> threadIdx.x) & 1u)) {
                    __threadfence_system();
                    *reinterpret_cast<volatile int32_t*>(c0.flag_signal[threadIdx.x]) = day + 1;
                }
            }
            __syncthreads();
        }
    }
}

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
    *reinterpret_cast<uchar4*>(c.ib_wr[0] + c.lo + a0) = make_uchar4(ibn[0], ibn[1], ibn[2], ibn[3]);
}

// peer mode: after a stand-alone k_snapshot, tell every GPU that this rank's bytes for `day` are complete
__global__ void k_signal(const DevCtx c, const int day)
{
    if (threadIdx.x < c.n_peers && ((c.signal_mask >> threadIdx.x) & 1u)) {
        __threadfence_system();
        *reinterpret_cast<volatile int32_t*>(c.flag_signal[threadIdx.x]) = day;
    }
}

}  // namespace dys
