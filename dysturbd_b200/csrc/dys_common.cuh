// Shared definitions of the sm_100a kernels: status sets, the device context (structure of arrays), counters.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "philox.cuh"

namespace dys {

// reference status code (agents[:,13], contagious3.py:115-118) times two; 0 pads the arrays to 1024 agents
enum : uint8_t { ST_PAD = 0, ST_S = 2, ST_I = 4, ST_QH = 6, ST_QI = 7, ST_D = 8, ST_H = 10, ST_R = 12, ST_X = 14 };

// status sets as bit masks over the code: membership is one shift
constexpr uint32_t sbit(int s) { return 1u << s; }
constexpr uint32_t M_SUS = sbit(ST_S) | sbit(ST_QH);                                  // `uninfected`, contagious3.py:197
constexpr uint32_t M_INF = sbit(ST_I) | sbit(ST_QI) | sbit(ST_D);                     // `infected`,   contagious3.py:196
constexpr uint32_t M_ISO = sbit(ST_QH) | sbit(ST_QI) | sbit(ST_D);                    // home slot only, :184
constexpr uint32_t M_ACTIVE = M_INF | sbit(ST_H);                                     // loop condition, :176
constexpr uint32_t M_EVER = M_ACTIVE | sbit(ST_R) | sbit(ST_X);                       // status not in {1, 3}
__device__ __forceinline__ bool in_set(uint32_t mask, uint32_t st) { return (mask >> st) & 1u; }

// day-start "infectious byte" of an agent (phase A, contagious3.py:181-202): bit7 infectious today (status 2 / 3.5 / 4
// at day start), bit6 isolated (3.5 / 4: only the home slot is visited), bits0-5 days since infection (pdf index)
constexpr uint8_t IB_INF = 0x80, IB_ISO = 0x40, IB_DAYS = 0x3F;
__device__ __forceinline__ uint8_t infectious_byte(uint32_t st, int day, int t_inf)
{
    if (!in_set(M_INF, st)) return 0;
    int n = day - t_inf;
    n = n < 0 ? 0 : (n > 63 ? 63 : n);
    return (uint8_t)(IB_INF | (st != ST_I ? IB_ISO : 0) | n);
}

// transposed interaction_prob entry word: bits0-26 row agent u (index local to the shard), bits28-31 the static
// shared-building class of the pair (u = row = the susceptible side, i = column = the infectious side)
constexpr uint32_t COL_MASK = 0x07FFFFFFu;
constexpr int CLS_SHIFT = 28;
constexpr uint32_t CLS_HH = 1u, CLS_HO = 2u, CLS_OH = 4u, CLS_OO = 8u;

constexpr int N_STATS = 64, N_SLOTS = 64, HIST0 = 32, HIST_LEN = 21;
enum {
    ST_ACTIVE = 0, ST_DAILY_INF, ST_RECOVERED, ST_QUARANTINED, ST_DAILY_QUAR, ST_HOSPITALIZED, ST_DAILY_HOSP, ST_DEAD,
    ST_DAILY_DEATHS, ST_EVER_INFECTED, ST_NEW_DIAG, ST_N_EVENTS, ST_SUS0, ST_INF0, ST_EDGES_SCANNED, ST_CANDIDATES,
    ST_EXPOSED, ST_HITS, ST_CHANGED
};
// per-day flag block (double buffered by day parity): [0] somebody is diagnosed on that day, [1] events claimed
constexpr int DF_ANY_ND = 0, DF_N_EVENTS = 1, DF_LEN = 4;

struct DevCtx {
    int64_t n_loc, n_pad, n_glob, lo;    // owned agents, padded count (multiple of 1024), global agents, first owned
    int64_t rt_lo, rt_n;                 // halo: global agents [rt_lo, rt_lo + rt_n) are the columns owned rows reference
    int64_t H, B, S;                     // OWNED households (hh[] is local to them), ALL buildings, ALL areas
    int64_t B_out, S_out;                // owned buildings / areas: home[] and sa[] are local to them, and so are the outputs
    int32_t V;
    int32_t recover, hospital_recover, diagnosis, quarantine, adm_min, adm_max, death_min;
    double norm_factor, compliance;
    uint64_t seed;
    // mutable state, owned agents
    uint8_t* status;
    int16_t *t_inf, *t_q, *t_adm;
    int32_t* src;                        // infector (global index), -1 = never infected
    // static per owned agent
    const int32_t *hh, *home, *sa;
    const double *risk, *p_adm, *p_death;
    const int32_t* routine;              // [rt_n * V] building index or -1 (halo agents: both sides of a pair are read)
    // interaction_prob TRANSPOSED over the halo columns: for infectious column i, the owned rows u with ip[u,i] != 0
    const int64_t* tptr;                 // [rt_n + 1]
    const uint32_t* tcolx;               // [nnz] u (local) | class << 28
    const double* tval;                  // [nnz] ip[u, i]
    const double* pdf;                   // [64]
    // day-start snapshot bytes of ALL agents, double buffered by day parity: k_push(d) reads ib_rd = buffer d&1, k_agent(d)
    // writes tomorrow's bytes of the owned agents into buffer (d+1)&1 -- of this GPU and, in peer mode, of every peer GPU
    // (P2P stores over NVLink), so the per-day exchange is fused into the producing kernel
    const uint8_t* ib_rd;
    uint8_t* ib_wr[8];                   // [n_wr] write targets (own buffer first)
    int64_t pub_lo[8], pub_n[8];         // peer mode: byte range of the owned slice that target p needs (its halo)
    int32_t n_wr;
    int32_t n_peers, rank;               // peer mode: n_peers > 1
    uint32_t wait_mask, signal_mask;     // ranks whose bytes this rank reads (waits for) / ranks that read this rank's bytes
    const int32_t* flag_wait;            // [n_peers] on THIS GPU: flag_wait[r] = last day rank r's bytes are complete for
    int32_t* flag_signal[8];             // [n_peers] &flags[rank] on every GPU (own included)
    const uint8_t* open;                 // [B]
    int32_t n_closed;                    // closed buildings today (counted on the host by dys_set_open)
    const uint8_t *hh_flag, *hh_qflag;   // [H] households diagnosed TODAY / quirk households (contagious3.py:263, :267)
    uint8_t *hh_flag_next, *hh_qflag_next;   // the other parity: tomorrow's, raised by k_agent, cleared by k_push
    const uint8_t* hh_always;            // [H] or null
    const int64_t* trig_ptr;             // [n_pad + 1] or null
    const int32_t* trig_hh;
    int32_t *df, *df_next;               // today's / tomorrow's flag block
    unsigned long long* part;            // [N_STATS][N_SLOTS] spread partial sums
    int64_t* stats;                      // [N_STATS]   (stats, sa_hist, bcount, ev live in one per-parity output block)
    unsigned int* ticket;
    int32_t* sa_hist;                    // [S * 21] or null
    int32_t* bcount;                     // [B] or null
    int2* ev;                            // [n_pad] today's (infected, infector) pairs, or null
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

// 16 bytes at a time for the byte arrays (the host pads them by 16)
__device__ __forceinline__ void grid_clear16(uint8_t* p, int64_t n_bytes)
{
    if (!p) return;
    uint4* q = reinterpret_cast<uint4*>(p);
    const int64_t n = (n_bytes + 15) >> 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        q[i] = make_uint4(0, 0, 0, 0);
}

// exact exposure test for days with closed buildings (contagious3.py:181-189, :230-231): a shared building must be
// open, and an isolated agent only keeps slot 0.  ug / ig index the halo-relative routine table.
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

}  // namespace dys
