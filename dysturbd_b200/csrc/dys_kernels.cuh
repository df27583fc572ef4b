// Hand-written sm_100a kernels for one day of DySTUrbD's epidemic step (contagious3.py:176-366).
//
//   k_day              phases B hospitalisation, C mortality, D pairwise infection, E transitions (contagious3.py:205-259)
//                      persistent CTAs; every tile's inputs arrive by TMA bulk copies (cp.async.bulk + mbarrier),
//                      double buffered, so the only global loads left in the loop are the infectious-byte gathers
//   k_household_stats  phase F household quarantine + G integer statistics (contagious3.py:261-366); also writes
//                      TOMORROW's day-start snapshot (phase A, :181-202), so a day is two launches
//   k_snapshot         phase A on its own (first day after a load / set_state)
//   k_edge_classes     load time: static shared-building class of every interaction_prob non-zero
//
// Everything here is HBM-bound integer/byte streaming: nothing is a dense contraction, so there is no tensor-core
// path.  Layout and roofline reasoning: DESIGN.md.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "philox.cuh"

namespace dys {

enum : uint8_t { ST_PAD = 0, ST_S = 2, ST_I = 4, ST_QH = 6, ST_QI = 7, ST_D = 8, ST_H = 10, ST_R = 12, ST_X = 14 };

// status sets as bit masks over the code: membership is one shift
constexpr uint32_t sbit(int s) { return 1u << s; }
constexpr uint32_t M_SUS = sbit(ST_S) | sbit(ST_QH);                                  // `uninfected`, contagious3.py:197
constexpr uint32_t M_INF = sbit(ST_I) | sbit(ST_QI) | sbit(ST_D);                     // `infected`,   contagious3.py:196
constexpr uint32_t M_ISO = sbit(ST_QH) | sbit(ST_QI) | sbit(ST_D);                    // home slot only, :184
constexpr uint32_t M_ACTIVE = M_INF | sbit(ST_H);                                     // loop condition, :176
constexpr uint32_t M_EVER = M_ACTIVE | sbit(ST_R) | sbit(ST_X);                       // status not in {1, 3}
__device__ __forceinline__ bool in_set(uint32_t mask, uint32_t st) { return (mask >> st) & 1u; }

// day-start "infectious byte" of an agent: bit7 infectious today (status 2 / 3.5 / 4 at day start), bit6 isolated
// (3.5 / 4: only the home slot is visited), bits0-5 days since infection (index into the pdf table)
constexpr uint8_t IB_INF = 0x80, IB_ISO = 0x40, IB_DAYS = 0x3F;
__device__ __forceinline__ uint8_t infectious_byte(uint32_t st, int day, int t_inf)
{
    if (!in_set(M_INF, st)) return 0;
    int n = day - t_inf;
    n = n < 0 ? 0 : (n > 63 ? 63 : n);
    return (uint8_t)(IB_INF | (st != ST_I ? IB_ISO : 0) | n);
}

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
// per-step flag block (double buffered by step parity): [0] a diagnosis happened today, [1] events claimed today
constexpr int DF_ANY_ND = 0, DF_N_EVENTS = 1, DF_LEN = 4;

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
    const int64_t* rowptr;               // [n_pad + 8]
    const uint32_t* colx;                // [nnz (+pad)] column | class << 28
    const double* val;                   // [nnz]
    const double* pdf;                   // [64]
    uint8_t* ib;                         // [n_glob padded] day-start infectious bytes of ALL agents
    const uint8_t* open;                 // [B]
    const int32_t* n_closed;             // [1] closed buildings today
    uint8_t *hh_flag, *hh_qflag;         // [H] this step's diagnosed households / quirk households
    uint8_t *hh_flag_next, *hh_qflag_next;   // the other parity: cleared by k_household_stats for the next step
    const uint8_t* hh_always;            // [H] or null
    const int64_t* trig_ptr;             // [n_pad + 1] or null
    const int32_t* trig_hh;
    int32_t *df, *df_next;               // this / next step's flag block
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

// block-level: fold per-thread counters into the spread partial-sum slots; warp sums by redux, one global atomic per
// counter per block, all issued by warp 0 so that a single fence orders them before the block's ticket
template <int NC>
__device__ __forceinline__ void block_accumulate(const DevCtx& c, const int (&idx)[NC], const unsigned int (&v)[NC],
                                                 unsigned int* s_acc /* [>= NC], zeroed, block-synchronised */)
{
    static_assert(NC <= 32, "one warp publishes the counters");
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const unsigned int x = __reduce_add_sync(0xffffffffu, v[k]);
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

// 16 bytes at a time for the byte arrays (sizes are padded to 16 by the host)
__device__ __forceinline__ void grid_clear16(uint8_t* p, int64_t n_bytes)
{
    if (!p) return;
    uint4* q = reinterpret_cast<uint4*>(p);
    const int64_t n = (n_bytes + 15) >> 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        q[i] = make_uint4(0, 0, 0, 0);
}

// ---- TMA bulk copy + mbarrier (PTX ISA 8.x; SASS: UBLKCP / SYNCS) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
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
// k_day.  Every WARP is an autonomous worker: it walks segments of 32 consecutive agents (lane = agent); a segment's
// interaction_prob column words are ONE contiguous run, fetched by a single TMA bulk copy into the warp's own
// double-buffered shared-memory stage (the copy of segment k+1 is in flight while segment k is processed; per-agent
// state of the next segment is prefetched into registers).  No block-level barrier exists in the loop.
//   scan   : the run is read in lock step (LDS.128), the day-start infectious byte of each column agent is gathered
//            (L1/L2 resident: neighbours live in the same community); infectious columns whose static class can
//            still expose are compacted into a per-warp candidate list (ballot + prefix popcount)
//   settle : the list is processed densely, one candidate per lane: row lookup (5-step binary search over the
//            segment's row offsets), exact exposure, P = ((ip*risk)*pdf)*norm in float64, keyed Philox Bernoulli;
//            the infector of a newly infected agent is the LARGEST infecting row index (contagious3.py:248) =
//            shared-memory atomicMax, order independent
//   agent  : B, C, E on the lane's own agent.
// ---------------------------------------------------------------------------------------------------------------
constexpr int WARPS = 8;                 // warps per CTA
constexpr int SEG = 32;                  // agents per segment
constexpr int CHW = 640;                 // column words staged per segment (2.5 KB); longer runs spill to direct loads
constexpr int CAP = 256;                 // candidate list entries per warp
constexpr uint32_t REL_MASK = 0x01FFFFFFu;   // candidate record: edge offset in the segment | days << 25 | iso << 31

struct __align__(128) WarpSmem {
    uint32_t col[2][CHW];
    uint32_t cand[CAP];
    int rel[SEG + 4];
    int best[SEG];
    uint64_t bar[2];
};
constexpr size_t K_DAY_SMEM = WARPS * sizeof(WarpSmem);

struct AgentRegs {                        // one lane's agent, prefetched a segment ahead
    uint32_t st;
    int tinf, tq, tadm;
};

__device__ __forceinline__ void load_agent(const DevCtx& c, int64_t a, AgentRegs& r)
{
    r.st = c.status[a];
    r.tinf = c.t_inf[a];
    r.tq = c.t_q[a];
    r.tadm = c.t_adm[a];
}

__device__ __forceinline__ void issue_segment(const DevCtx& c, int64_t e0, int64_t e1, uint32_t* dst, uint64_t* bar)
{
    const int64_t al0 = e0 & ~3LL;
    int64_t words = ((e1 - al0) + 3) & ~3LL;
    if (words > CHW) words = CHW;
    if (words > 0) {
        mbar_expect_tx(bar, (uint32_t)words * 4u);
        bulk_g2s(dst, c.colx + al0, (uint32_t)words * 4u, bar);
    } else {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
}

__global__ void __launch_bounds__(WARPS * 32, 4) k_day(const DevCtx c, const int day)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double s_pdf[64];
    __shared__ unsigned int s_acc[16];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    WarpSmem& ws = reinterpret_cast<WarpSmem*>(smem_raw)[wid];
    const int64_t n_seg = c.n_pad / SEG;
    const int64_t stride = (int64_t)gridDim.x * WARPS;

    // scratch of k_household_stats (nothing else touches it while this kernel runs)
    grid_clear(c.sa_hist, c.S * HIST_LEN);
    grid_clear(c.bcount, c.B);
    grid_clear(c.io_mat, c.io_mat ? c.S * c.S : 0);

    if (lane == 0) {
        mbar_init(&ws.bar[0], 1);
        mbar_init(&ws.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 64) s_pdf[tid] = c.pdf[tid];
    if (tid < 16) s_acc[tid] = 0;
    __syncthreads();
    const bool any_closed = *c.n_closed != 0;

    // sus0, inf0, admissions, deaths, scanned, candidates, exposed, hits, events, new diagnoses, changed
    unsigned int cnt[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

    int64_t seg = (int64_t)blockIdx.x * WARPS + wid;
    AgentRegs cur{}, nxt{};
    int64_t rp_cur = 0, rp_cur_hi = 0, rp_nxt = 0, rp_nxt_hi = 0;
    if (seg < n_seg) {
        const int64_t a = seg * SEG + lane;
        rp_cur = c.rowptr[a]; rp_cur_hi = c.rowptr[a + 1];
        load_agent(c, a, cur);
        if (seg + stride < n_seg) {
            const int64_t an = (seg + stride) * SEG + lane;
            rp_nxt = c.rowptr[an]; rp_nxt_hi = c.rowptr[an + 1];
            load_agent(c, an, nxt);
        }
        const int64_t e0 = __shfl_sync(0xffffffffu, rp_cur, 0), e1 = __shfl_sync(0xffffffffu, rp_cur_hi, 31);
        if (lane == 0) issue_segment(c, e0, e1, ws.col[0], &ws.bar[0]);
    }

    for (int it = 0; seg < n_seg; seg += stride, ++it) {
        const int sidx = it & 1;
        const int64_t seg0 = seg * SEG, a = seg0 + lane;
        const bool have_next = seg + stride < n_seg;
        // ---- pipeline: copy of the next segment, registers of the one after ----
        AgentRegs nn{};
        int64_t rp_nn = 0, rp_nn_hi = 0;
        if (have_next) {
            const int64_t ne0 = __shfl_sync(0xffffffffu, rp_nxt, 0), ne1 = __shfl_sync(0xffffffffu, rp_nxt_hi, 31);
            if (lane == 0) issue_segment(c, ne0, ne1, ws.col[sidx ^ 1], &ws.bar[sidx ^ 1]);
            if (seg + 2 * stride < n_seg) {
                const int64_t ann = (seg + 2 * stride) * SEG + lane;
                rp_nn = c.rowptr[ann]; rp_nn_hi = c.rowptr[ann + 1];
                load_agent(c, ann, nn);
            }
        }
        const int64_t e0 = __shfl_sync(0xffffffffu, rp_cur, 0);
        const int T = (int)(__shfl_sync(0xffffffffu, rp_cur_hi, 31) - e0);
        const int off = (int)(e0 & 3);                 // word w of the staged run is edge (w - off) of the segment
        const uint32_t st0 = cur.st;
        const bool sus = in_set(M_SUS, st0);
        ws.rel[lane] = (int)(rp_cur - e0);
        if (lane == 0) ws.rel[SEG] = T;
        ws.best[lane] = -1;
        cnt[0] += sus;
        cnt[4] += sus ? (unsigned)(rp_cur_hi - rp_cur) : 0u;
        const unsigned sus_rows = __ballot_sync(0xffffffffu, sus);     // also orders the shared writes above
        mbar_wait(&ws.bar[sidx], (uint32_t)((it >> 1) & 1));

        if (sus_rows && T > 0) {
            const uint32_t* col = ws.col[sidx];
            int n = 0;
            auto settle = [&]() {
                for (int k0 = 0; k0 < n; k0 += 32) {
                    const int k = k0 + lane;
                    const bool valid = k < n;
                    const uint32_t rec = valid ? ws.cand[k] : 0u;
                    const int rel = (int)(rec & REL_MASK);
                    int lo = 0, hi = SEG;                      // ws.rel[lo] <= rel < ws.rel[hi]
#pragma unroll
                    for (int q = 0; q < 5; ++q) {
                        const int mid = (lo + hi) >> 1;
                        if (ws.rel[mid] <= rel) lo = mid; else hi = mid;
                    }
                    const uint32_t st_r = __shfl_sync(0xffffffffu, st0, lo);
                    if (!valid || !in_set(M_SUS, st_r)) continue;
                    ++cnt[5];
                    const int w = rel + off;
                    const uint32_t cw = w < CHW ? col[w] : __ldg(c.colx + e0 + rel);
                    const uint32_t ig = cw & COL_MASK, cls = cw >> CLS_SHIFT;
                    const bool iso_u = (st_r == ST_QH), iso_i = rec >> 31;
                    const int64_t ul = seg0 + lo, ug = c.lo + ul;
                    bool ex = (cls & CLS_HH) | ((cls & CLS_HO) && !iso_i) | ((cls & CLS_OH) && !iso_u) |
                              ((cls & CLS_OO) && !iso_u && !iso_i);
                    if (ex && any_closed)
                        ex = exposed_slow(c.routine, c.open, c.V, ug - c.rt_lo, (int64_t)ig - c.rt_lo, iso_u, iso_i);
                    if (!ex) continue;
                    ++cnt[6];
                    // contagious3.py:232-238, same association, separately rounded (no FMA can form: no adds)
                    double p = __dmul_rn(__ldg(c.val + e0 + rel), __ldg(c.risk + ul));
                    p = __dmul_rn(p, s_pdf[(rec >> 25) & 0x3Fu]);
                    p = __dmul_rn(p, c.norm_factor);
                    if (c.pair_prob) c.pair_prob[(size_t)ug * c.n_glob + ig] = p;
                    const double u = draw(c.u_pair, ug * c.n_glob + ig, c.seed, day, STREAM_INFECTION, (uint32_t)ug, ig);
                    if (p > u) { ++cnt[7]; atomicMax(&ws.best[lo], (int)ig); }        // :240, :248
                }
                __syncwarp();
                n = 0;
            };
            auto scan4 = [&](const uint32_t (&cw)[4], const uint32_t (&ibv)[4], const int rel0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    // static pre-filter: classes that need a non-isolated column are dead if the column is isolated
                    const uint32_t cls = cw[j] >> CLS_SHIFT, iso_i = (ibv[j] >> 6) & 1u;
                    const bool cnd = (ibv[j] & IB_INF) && (cls & (iso_i ? (CLS_HH | CLS_OH) : 0xFu));
                    const unsigned m = __ballot_sync(0xffffffffu, cnd);
                    if (cnd) ws.cand[n + __popc(m & lt_mask)] = (uint32_t)(rel0 + j) | ((ibv[j] & 0x7Fu) << 25);
                    n += __popc(m);
                }
            };
            const int wend = off + T;                          // one past the last valid staged word
            const int wstaged = wend < CHW ? wend : CHW;
            for (int w0 = 0; w0 < wstaged; w0 += 4 * 32) {
                if (n > CAP - 4 * 32) settle();
                const int w = w0 + 4 * lane;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (w < wstaged) v = *reinterpret_cast<const uint4*>(&col[w]);
                const uint32_t cw[4] = {v.x, v.y, v.z, v.w};
                uint32_t ibv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    ibv[j] = (w + j >= off && w + j < wend) ? (uint32_t)__ldg(c.ib + (cw[j] & COL_MASK)) : 0u;
                scan4(cw, ibv, w - off);
            }
            // runs longer than the staged window (fat rows): the rest straight from global memory
            for (int64_t w0 = CHW; w0 < wend; w0 += 4 * 32) {
                if (n > CAP - 4 * 32) settle();
                const int64_t w = w0 + 4 * lane;
                uint32_t cw[4] = {0, 0, 0, 0}, ibv[4] = {0, 0, 0, 0};
                if (w < wend) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4*>(c.colx + (e0 - off) + w));
                    cw[0] = v.x; cw[1] = v.y; cw[2] = v.z; cw[3] = v.w;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (w + j < wend) ibv[j] = __ldg(c.ib + (cw[j] & COL_MASK));
                }
                scan4(cw, ibv, (int)(w - off));
            }
            if (n > 0) settle();
        }
        __syncwarp();

        // ---- the lane's own agent: B, C on the day-start status, the infection found above, then the rules of E ----
        const int best = ws.best[lane];
        uint32_t s = st0;
        int tinf = cur.tinf, tq = cur.tq;
        bool w_tinf = false, w_tq = false, w_tadm = false;
        if (in_set(M_INF, st0)) {
            ++cnt[1];
            if (tinf >= c.adm_min && tinf <= c.adm_max) {                   // contagious3.py:205 (absolute infection day)
                const double u = draw(c.u_adm, c.lo + a, c.seed, day, STREAM_ADMISSION, (uint32_t)(c.lo + a), 0);
                if (__ldg(c.p_adm + a) > u) { s = ST_H; c.t_adm[a] = (int16_t)day; ++cnt[2]; }   // :213-216
            }
        } else if (st0 == ST_H && cur.tadm >= c.death_min) {               // :219
            const double u = draw(c.u_death, c.lo + a, c.seed, day, STREAM_DEATH, (uint32_t)(c.lo + a), 0);
            if (__ldg(c.p_death + a) > u) { s = ST_X; w_tadm = true; ++cnt[3]; }                 // :225-227, :259
        }
        const bool infected_today = best >= 0;
        if (infected_today) {                                               // :241-248
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
        if (s != st0) c.status[a] = (uint8_t)s;
        if (new_diag) {
            c.hh_flag[c.hh[a]] = 1;                   // same-value stores: race-free in effect
            if (c.trig_ptr)
                for (int64_t t = c.trig_ptr[a]; t < c.trig_ptr[a + 1]; ++t) c.hh_qflag[c.trig_hh[t]] = 1;
            c.df[DF_ANY_ND] = 1;
        }
        cnt[8] += infected_today;
        cnt[9] += new_diag;
        cnt[10] += (s != st0) | w_tq;

        // today's infection events: one slot claim per warp that has any (order is arbitrary; hosts sort)
        if (c.ev_agent) {
            const unsigned evm = __ballot_sync(0xffffffffu, infected_today);
            if (evm) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&c.df[DF_N_EVENTS], __popc(evm));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (infected_today) {
                    const int slot = base + __popc(evm & lt_mask);
                    c.ev_agent[slot] = (int)(c.lo + a);
                    c.ev_src[slot] = best;
                }
            }
        }
        __syncwarp();      // stage sidx, rel[] and best[] are free again
        cur = nxt; nxt = nn;
        rp_cur = rp_nxt; rp_cur_hi = rp_nxt_hi; rp_nxt = rp_nn; rp_nxt_hi = rp_nn_hi;
    }
    const int idx[11] = {ST_SUS0, ST_INF0, ST_DAILY_HOSP, ST_DAILY_DEATHS, ST_EDGES_SCANNED, ST_CANDIDATES, ST_EXPOSED,
                         ST_HITS, ST_N_EVENTS, ST_NEW_DIAG, ST_CHANGED};
    __syncthreads();
    block_accumulate<11>(c, idx, cnt, s_acc);
}

// ---------------------------------------------------------------------------------------------------------------
// k_household_stats: phase F on the household flags raised by k_day, the integer statistics of the end-of-day
// state (phase G), and tomorrow's day-start snapshot byte (phase A).  4 agents per thread, vector loads.  The last
// block to finish folds the spread partial sums into the stats block.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_household_stats(const DevCtx c, const int day)
{
    __shared__ unsigned int s_hist[HIST_LEN];
    __shared__ unsigned int s_acc[16];
    __shared__ bool s_last;
    if (threadIdx.x < HIST_LEN) s_hist[threadIdx.x] = 0;
    if (threadIdx.x < 16) s_acc[threadIdx.x] = 0;
    // the other parity's flags: last read by yesterday's launch of this kernel, next written by tomorrow's k_day
    grid_clear16(c.hh_flag_next, c.H);
    grid_clear16(c.hh_qflag_next, c.H);
    if (blockIdx.x == 0 && threadIdx.x < DF_LEN) c.df_next[threadIdx.x] = 0;
    __syncthreads();
    const bool any_nd = c.df[DF_ANY_ND] != 0;
    const int64_t a0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    // active, daily_inf, recovered, quarantined, daily_quar, hosp, dead, ever, changed
    unsigned int cnt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (a0 < c.n_pad) {
        const uchar4 sv = *reinterpret_cast<const uchar4*>(c.status + a0);
        uint32_t s[4] = {sv.x, sv.y, sv.z, sv.w};
        uint32_t newq = 0;
        if (any_nd) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (s[j] != ST_S && s[j] != ST_I) continue;
                const int64_t a = a0 + j;
                const int h = c.hh[a];
                bool hit = c.hh_flag[h];                                              // contagious3.py:263
                if (s[j] == ST_I && !hit)                                             // :267 (whole-row isin)
                    hit = c.hh_qflag[h] || (c.hh_always && c.hh_always[h]);
                if (hit && c.compliance < 1.0)
                    hit = keyed_uniform(c.seed, day, STREAM_COMPLIANCE, (uint32_t)(c.lo + a), 0) < c.compliance;
                if (hit) {
                    s[j] = (s[j] == ST_S) ? ST_QH : ST_QI;                            // :265, :269
                    c.t_q[a] = (int16_t)day;                                          // :266, :270
                    newq |= 1u << j;
                }
            }
            if (newq) *reinterpret_cast<uchar4*>(c.status + a0) = make_uchar4(s[0], s[1], s[2], s[3]);
            cnt[8] += __popc(newq);
        }
        uint32_t m_ever = 0, m_q = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            m_ever |= in_set(M_EVER, s[j]) << j;
            m_q |= in_set(M_ISO, s[j]) << j;
        }
        short4 ti4 = make_short4(0, 0, 0, 0), tq4 = make_short4(0, 0, 0, 0);
        if (m_ever) ti4 = *reinterpret_cast<const short4*>(c.t_inf + a0);
        if (m_q & ~newq) tq4 = *reinterpret_cast<const short4*>(c.t_q + a0);
        const int ti[4] = {ti4.x, ti4.y, ti4.z, ti4.w}, tq[4] = {tq4.x, tq4.y, tq4.z, tq4.w};
        uint8_t ibn[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t st = s[j];
            ibn[j] = infectious_byte(st, day + 1, ti[j]);                  // tomorrow's phase A (contagious3.py:192-202)
            cnt[0] += in_set(M_ACTIVE, st);
            cnt[2] += (st == ST_R);
            cnt[3] += (m_q >> j) & 1u;
            cnt[4] += ((m_q >> j) & 1u) && (((newq >> j) & 1u) || tq[j] == day);
            cnt[5] += (st == ST_H);
            cnt[6] += (st == ST_X);
            if ((m_ever >> j) & 1u) {
                const int64_t a = a0 + j;
                ++cnt[7];
                const int k = day - ti[j];
                cnt[1] += (k == 0);
                if (k >= 0 && k < HIST_LEN) {
                    atomicAdd(&s_hist[k], 1u);
                    if (c.sa_hist) { const int sa = c.sa[a]; if (sa >= 0) atomicAdd(&c.sa_hist[(size_t)sa * HIST_LEN + k], 1); }
                }
                if (c.bcount && in_set(M_INF, st)) atomicAdd(&c.bcount[c.home[a]], 1);
                if (c.io_mat && k == 0) {
                    // contagious3.py:356-366: row = area of today's infected, column = area of the infector
                    const int srci = c.src[a];
                    if (srci >= 0) {
                        const int sa_u = c.sa[a], sa_i = c.sa[srci - c.lo];
                        if (sa_u >= 0 && sa_i >= 0) atomicAdd(&c.io_mat[(size_t)sa_u * c.S + sa_i], 1);
                    }
                }
            }
        }
        *reinterpret_cast<uchar4*>(c.ib + c.lo + a0) = make_uchar4(ibn[0], ibn[1], ibn[2], ibn[3]);
    }
    const int idx[9] = {ST_ACTIVE, ST_DAILY_INF, ST_RECOVERED, ST_QUARANTINED, ST_DAILY_QUAR, ST_HOSPITALIZED, ST_DEAD,
                        ST_EVER_INFECTED, ST_CHANGED};
    __syncthreads();
    block_accumulate<9>(c, idx, cnt, s_acc);
    if (threadIdx.x < 32) {
        // warp 0 published the counters above; it also publishes the histogram, fences, and takes the ticket
        if (threadIdx.x < HIST_LEN && s_hist[threadIdx.x])
            atomicAdd(&c.part[(size_t)(HIST0 + threadIdx.x) * N_SLOTS + (blockIdx.x & (N_SLOTS - 1))],
                      (unsigned long long)s_hist[threadIdx.x]);
        __threadfence();
        __syncwarp();
        if (threadIdx.x == 0) s_last = (atomicAdd(c.ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        // part[counter][slot]: a warp folds one counter with two coalesced loads and a redux
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

// phase A alone: the day-start snapshot byte of every owned agent (first day after a load / dys_set_state)
__global__ void __launch_bounds__(256) k_snapshot(const DevCtx c, const int day)
{
    const int64_t a0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    if (a0 >= c.n_pad) return;
    const uchar4 sv = *reinterpret_cast<const uchar4*>(c.status + a0);
    const short4 ti = *reinterpret_cast<const short4*>(c.t_inf + a0);
    *reinterpret_cast<uchar4*>(c.ib + c.lo + a0) =
        make_uchar4(infectious_byte(sv.x, day, ti.x), infectious_byte(sv.y, day, ti.y), infectious_byte(sv.z, day, ti.z),
                    infectious_byte(sv.w, day, ti.w));
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
