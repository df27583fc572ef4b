// Shared definitions of the sm_100a kernels: status sets, the device context (structure of arrays), counters.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "philox.cuh"

// DYS_CHECKS=1 (the "checked" build, dysturbd_b200/_build.py --checked): bounds checks on every index the day kernels form
// -- work list, frontier, edge records, mailboxes, member lists, output blocks, remote snapshot bytes, work words.  A
// failed check stores 1000 + its code in the abort word; the launch ends at its next spin check and the host reports
// DYS_ERR_ABORTED with the code.  (compute-sanitizer is not available on the pool these kernels are developed on.)
#ifndef DYS_CHECKS
#define DYS_CHECKS 0
#endif
#if DYS_CHECKS
#define DYS_ASSERT(ctx, cond, code) do { if (!(cond) && (ctx).abort_word) { *(ctx).abort_word = 1000 + (code); *(ctx).abort_host = 1000 + (code); } } while (0)
#else
#define DYS_ASSERT(ctx, cond, code) do { } while (0)
#endif

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
constexpr uint32_t EDGE_ONE_BLD = 0x80000000u;      // pad: the pair shares exactly one building (low 31 bits)
constexpr uint32_t EDGE_MASKS = 0x40000000u;        // pad: bits 0-9 slots of u that meet any slot of i, bits 10-19 ... the home slot of i
// one non-zero of the transposed interaction_prob, 16 bytes so that a lane fetches it with one 128-bit load.  The first
// product of contagious3.py:232 -- interaction_prob[u,i] * agents[u,14] -- only involves static data, so it is formed once
// at load time (same IEEE multiply, same rounding) and the push never gathers risk[u].
struct __align__(16) TEdge {
    uint32_t colx;                       // u | class << 28
    uint32_t pad;                        // EDGE_ONE_BLD | b: the pair shares exactly one building, b (closures: exposure = class && open[b])
    double val;                          // ip[u,i] * risk[u]
};

constexpr int N_STATS = 64, N_SLOTS = 64, HIST0 = 32, HIST_LEN = 21;
enum {
    ST_ACTIVE = 0, ST_DAILY_INF, ST_RECOVERED, ST_QUARANTINED, ST_DAILY_QUAR, ST_HOSPITALIZED, ST_DAILY_HOSP, ST_DEAD,
    ST_DAILY_DEATHS, ST_EVER_INFECTED, ST_NEW_DIAG, ST_N_EVENTS, ST_SUS0, ST_INF0, ST_EDGES_SCANNED, ST_CANDIDATES,
    ST_EXPOSED, ST_HITS, ST_CHANGED
};
// per-day flag block, a ring of four indexed by day & 3: [0] somebody is diagnosed on that day, [1] events claimed
constexpr int DF_ANY_ND = 0, DF_N_EVENTS = 1, DF_LEN = 4, DF_RING = 4;
constexpr int Q_RING = 3, PART_RING = 3;         // quirk household flags and partial sums: rings indexed by day % 3
// per-agent attention byte (double buffered by day parity): bit0 "k_push left an infector in your mailbox",
// bit1 "somebody of your household is diagnosed today" (contagious3.py:263).  An agent whose byte is 0 and who is
// susceptible (or long recovered) cannot change today: the sweep of agent_body reads two bytes for it and moves on.
constexpr uint32_t ATT_HIT = 1u, ATT_HH = 2u;
constexpr int TRACE_SLOTS = 12;                  // developer tracing (DYS_TRACE): timestamps per block and day

// Everything the settle phase needs about one agent besides its status byte, in ONE 32-byte record = one memory sector:
// a gather costs a round trip per dependent access and a sector per distinct column, so the agent's days in state, its
// area, home building, household (with where its members are listed) and admission probability arrive together.
struct __align__(16) AgentRec {
    int16_t t_inf, t_q, t_adm;           // agents[:,16], agents[:,19], agents[:,24] (contagious3.py:115-118); 0 = never
    int16_t hh_size;                     // members of the agent's household ...
    int32_t sa;                          // statistical area (local to the owned areas), -1 = none
    int32_t home;                        // home building (local to the owned buildings)
    int32_t hh;                          // household (local to the owned households)
    int32_t hh_slot;                     // ... listed at hh_mem[hh_slot .. hh_slot + hh_size)
    double p_adm;                        // agents[:,23]
};
static_assert(sizeof(AgentRec) == 32, "one sector");

// Monte-Carlo replicas of one loaded city (BASELINE.json configs[4]): the parameters that differ between them
struct RepParam {
    double norm_factor, compliance;
    uint64_t seed;
};

struct DevCtx {
    int64_t n_loc, n_pad, n_glob, lo;    // owned agents, padded count (multiple of 1024), global agents, first owned
    int64_t rt_lo, rt_n;                 // halo: global agents [rt_lo, rt_lo + rt_n) are the columns owned rows reference
    int64_t H, B, S;                     // OWNED households (hh[] is local to them), ALL buildings, ALL areas
    int64_t B_out, S_out;                // owned buildings / areas: home[] and sa[] are local to them, and so are the outputs
    int32_t V;
    int32_t recover, hospital_recover, diagnosis, quarantine, adm_min, adm_max, death_min;
    double norm_factor, compliance;
    uint64_t seed;
    int32_t r_idle_day;                  // from this day on every recovered agent has left the 21-bin histogram (see k_validate_population)
    // Replicas: n_rep independent epidemics on the SAME city in one launch.  Everything mutable (and the per-day buffers and
    // output blocks) exists n_rep times, replica r at `r * stride` behind replica 0; the network, routines, households and
    // probabilities are shared.  shift_replica() turns a replica-0 view into replica r's.
    int32_t n_rep, rep;
    int64_t ib_stride, out_stride;       // bytes between two replicas' snapshot buffers / output blocks
    const RepParam* rep_params;          // [n_rep] or null (n_rep == 1: the fields above)
    // spinning loops (grid barrier, peer waits) give up when *abort_word != 0 or after spin_timeout_ns (then they set it)
    volatile int32_t* abort_word;        // DEVICE memory: polled by the spinning loops (the host sets it with a copy on a side stream)
    volatile int32_t* abort_host;        // its mirror in mapped pinned host memory: written (never polled) by the device when a
                                         // wait times out or a check fails, read by the host without any CUDA call
    unsigned long long spin_timeout_ns;
    // mutable state, owned agents
    uint8_t* status;
    AgentRec* rec;                       // days in state (mutable) + area and home building (static)
    int32_t* src;                        // infector (global index), -1 = never infected
    // static per owned agent
    const int32_t* hh;
    const double *risk, *p_adm, *p_death;
    const int32_t* routine;              // [rt_n * V] building index or -1 (halo agents: both sides of a pair are read)
    const int64_t* hh_ptr;               // [H + 1] household -> its members (local agent indices)
    const int32_t* hh_mem;               // [n_loc]
    // interaction_prob TRANSPOSED over the halo columns: for infectious column i, the owned rows u with ip[u,i] != 0
    const int64_t* tptr;                 // [rt_n + 1]
    const TEdge* tedge;                  // [nnz]
    const double* pdf;                   // [64]
    const uint8_t* hh_always;            // [H] or null   (quirk of contagious3.py:267, see dys_load_quirk)
    const int64_t* trig_ptr;             // [n_pad + 1] or null
    const int32_t* trig_hh;
    // peer mode (n_peers > 1): every rank stores tomorrow's snapshot bytes straight into the buffers of the ranks whose halo
    // contains them and counts what it delivered in a counter on the receiving GPU: one unit per 128-agent piece plus one
    // token per day, so that partners stay within one day of each other even when no bytes travel in one direction
    int64_t pub_lo[8], pub_n[8];         // write target p >= 1: GLOBAL agents [pub_lo, pub_lo + pub_n) of the owned slice lie in its halo
    int32_t n_wr;
    int32_t n_peers, rank;
    const unsigned long long* arrive;    // [n_peers] on THIS GPU: units received from rank r so far
    unsigned long long* arrive_at[8];    // write target p >= 1: the counter of this rank on that GPU
    uint32_t expect_in[8];               // units per day from rank r (0: not a partner)
    uint32_t units_out[8];               // write target p >= 1: units per day this rank delivers there (pieces + 1)
    // the owned pieces that lie in SOME partner's halo form a prefix [0, bnd_a) and a suffix [bnd_b, n_pieces) of the owned
    // range (shards are contiguous runs of communities): k_days settles them FIRST every day, spread over all blocks, so
    // that their bytes reach the partners while the bulk of the day is still being processed (bnd_ok = 0: no such split)
    int32_t bnd_ok, bnd_a, bnd_b;
    unsigned long long epoch;            // publications that preceded this day's own: the bytes this day READS were the epoch-th
    // ---- the view of ONE day (bind_day) ----
    // day-start snapshot bytes of ALL agents, double buffered by day parity: the scan push of day d reads buffer d&1, the
    // agent pass of day d writes tomorrow's bytes of the owned agents into buffer (d+1)&1
    const uint8_t* ib_rd;
    uint8_t* ib_wr[8];                   // [n_wr] write targets (own buffer first; peers' buffers are written by publish_body)
    int32_t write_ib;                    // 0: nobody will scan tomorrow's bytes (a middle day of a single-GPU k_days window)
    const uint8_t* open;                 // [B]
    int32_t n_closed;                    // closed buildings today (counted on the host, or by the device policy)
    int64_t open_stride;                 // device lockdown policy: every replica has its own flag vector, B bytes apart (else 0)
    const int32_t* n_closed_ptr;         // ... and its own count (else null)
    uint8_t *att, *att_next;             // [n_pad] attention bytes of this day / of tomorrow
    int32_t *mb_any, *mb_iso;            // [n_pad] this day's mailboxes: largest infecting i if u is free / if u is isolated
    uint8_t *qflag, *qflag_next;         // [H] quirk households of this day / tomorrow (contagious3.py:267)
    int32_t *df, *df_next;               // this day's / tomorrow's flag block
    unsigned long long* part;            // [N_STATS][N_SLOTS] spread partial sums of this day
    int64_t* stats;                      // [N_STATS]   (stats, sa_hist, bcount, ev live in one output block per day)
    unsigned int* ticket;
    int32_t* sa_hist;                    // [S_out * 21] or null
    int32_t* bcount;                     // [B_out] or null
    int2* ev;                            // [n_pad] today's (infected, infector) pairs, or null
    int32_t* io_mat;                     // [S * S] or null
    // k_days only: the agents are dealt to the blocks in contiguous runs of 128-agent pieces, re-cut every day so that
    // every block gets the same MEASURED work (the epidemic is never evenly spread over the population)
    int32_t* bounds_plan;                // [n_work + 1] where the cut made from today's costs goes (used in three days), or null
    int32_t cut_first, cut_last;         // this block's run of pieces today, or cut_first < 0: an equal share
    uint32_t* piece_work;                // [n_pad / 128] today's cost of each piece, written by the sweep, or null
    const double *u_adm, *u_death, *u_pair;   // injected draws or null
    double* pair_prob;                   // [n_glob^2] or null (tests)
};

// the buffers behind the per-day view; bind_day() points a DevCtx at the ones of `day`
struct DayBuffers {
    uint8_t* att[2];
    int32_t *mb_any[2], *mb_iso[2];
    uint8_t* qflag[Q_RING];
    int32_t* df;                         // [DF_RING][DF_LEN]
    unsigned long long* part;            // [PART_RING][N_STATS][N_SLOTS]
    uint8_t* ib[2];                      // own snapshot buffers by day parity
    uint8_t* peer_ib[8][2];              // peer mode: every rank's, by day parity
    int32_t* bounds[3];                  // k_days: the cut of day d is bounds[d % 3], made from the costs of day d - 3 or
                                         // earlier -- complete a whole day before it is used, so that a block can fetch its
                                         // run while it builds tomorrow's view; or null
    uint32_t* piece_work[2];             // k_days: the piece costs measured on day d are piece_work[d & 1]
};

// where snapshot bytes of parity `parity` are written: the own buffer, then (peer mode) every other rank's
__host__ __device__ __forceinline__ void bind_ib_targets(DevCtx& c, const DayBuffers& b, const int parity)
{
    c.n_wr = 0;
    c.ib_wr[c.n_wr++] = b.ib[parity];
    for (int r = 0; r < c.n_peers; ++r)      // (partners only, in rank order: the order of pub_lo / arrive_at / units_out)
        if (c.n_peers > 1 && r != c.rank && c.expect_in[r]) c.ib_wr[c.n_wr++] = b.peer_ib[r][parity];
}

__host__ __device__ __forceinline__ void bind_day(DevCtx& c, const DayBuffers& b, const int day)
{
    const int p = day & 1;
    c.att = b.att[p]; c.att_next = b.att[p ^ 1];
    c.mb_any = b.mb_any[p]; c.mb_iso = b.mb_iso[p];
    const int R = c.n_rep > 1 ? c.n_rep : 1;
    c.qflag = b.qflag[day % Q_RING]; c.qflag_next = b.qflag[(day + 1) % Q_RING];
    c.df = b.df + (size_t)(day & (DF_RING - 1)) * R * DF_LEN; c.df_next = b.df + (size_t)((day + 1) & (DF_RING - 1)) * R * DF_LEN;
    c.part = b.part + (size_t)(day % PART_RING) * R * N_STATS * N_SLOTS;
    c.bounds_plan = b.bounds[day % 3]; c.piece_work = b.piece_work[p];      // (day + 3) % 3
    c.cut_first = c.cut_last = -1;
    c.ib_rd = b.ib[p];
    c.write_ib = 1;
    bind_ib_targets(c, b, p ^ 1);        // the agent pass of `day` writes tomorrow's snapshot bytes
}

// replica 0's view of a day -> replica r's (every per-replica array is [n_rep][stride])
__host__ __device__ __forceinline__ void shift_replica(DevCtx& c, const int r)
{
    c.rep = r;
    if (c.n_rep <= 1) return;
    const size_t n = (size_t)c.n_pad * r;
    c.status += n; c.rec += n; c.src += n;
    c.att += (size_t)(c.n_pad + 16) * r; c.att_next += (size_t)(c.n_pad + 16) * r;
    c.mb_any += n; c.mb_iso += n;
    c.qflag += (size_t)(c.H + 16) * r; c.qflag_next += (size_t)(c.H + 16) * r;
    c.df += DF_LEN * r; c.df_next += DF_LEN * r;
    c.part += (size_t)N_STATS * N_SLOTS * r;
    c.ib_rd += (size_t)c.ib_stride * r; c.ib_wr[0] += (size_t)c.ib_stride * r;
    const size_t o = (size_t)c.out_stride * r;
    if (c.stats) c.stats = reinterpret_cast<int64_t*>(reinterpret_cast<unsigned char*>(c.stats) + o);
    if (c.sa_hist) c.sa_hist = reinterpret_cast<int32_t*>(reinterpret_cast<unsigned char*>(c.sa_hist) + o);
    if (c.bcount) c.bcount = reinterpret_cast<int32_t*>(reinterpret_cast<unsigned char*>(c.bcount) + o);
    if (c.ev) c.ev = reinterpret_cast<int2*>(reinterpret_cast<unsigned char*>(c.ev) + o);
    if (c.piece_work) c.piece_work += (size_t)(c.n_pad >> 7) * r;
    c.open += (size_t)c.open_stride * r;
#ifdef __CUDA_ARCH__
    if (c.n_closed_ptr) { c.n_closed_ptr += r; c.n_closed = __ldcg(c.n_closed_ptr); }
    if (c.rep_params) {
        const RepParam p = c.rep_params[r];
        c.norm_factor = p.norm_factor; c.compliance = p.compliance; c.seed = p.seed;
    }
#endif
}

__device__ __forceinline__ double draw(const double* inj, int64_t idx, uint64_t seed, uint32_t day, uint32_t stream,
                                       uint32_t a, uint32_t b)
{
    return inj ? __ldg(inj + idx) : keyed_uniform(seed, day, stream, a, b);
}

// block-level: fold per-thread counters into the spread partial-sum slots; warp sums by redux, one global atomic per
// counter per block, all issued by warp 0 so that a single fence orders them before the block's ticket
template <int NC>
__device__ __forceinline__ void block_accumulate(unsigned long long* part, const int (&idx)[NC], const unsigned int (&v)[NC],
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
        atomicAdd(&part[(size_t)idx[threadIdx.x] * N_SLOTS + (blockIdx.x & (N_SLOTS - 1))],
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

// exposure under closures from the edge's slot masks: some allowed slot of u whose building is open (u isolated: slot 0
// only; i isolated: only the slots that meet i's home slot) -- a load per set bit instead of the V x V walk below
__device__ __forceinline__ bool exposed_masks(const int32_t* __restrict__ routine, const uint8_t* __restrict__ open, int V, int64_t ug,
                                              uint32_t pad, bool iso_u, bool iso_i)
{
    uint32_t m = iso_i ? (pad >> 10) & 0x3FFu : pad & 0x3FFu;
    if (iso_u) m &= 1u;
    const int32_t* ru = routine + ug * V;
    while (m) {
        const int s = __ffs(m) - 1;
        m &= m - 1;
        if (__ldcg(open + __ldg(ru + s))) return true;
    }
    return false;
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
        if (bu < 0 || (s > 0 && iso_u) || !__ldcg(open + bu)) continue;      // (ld.cg: the device lockdown policy rewrites the flags inside a launch)
        for (int t = 0; t < V; ++t) {
            if (t > 0 && iso_i) break;
            if (__ldg(ri + t) == bu) return true;
        }
    }
    return false;
}

}  // namespace dys
