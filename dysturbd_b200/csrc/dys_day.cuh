// The kernels of a simulated day (contagious3.py:176-366).
//
//   push   phase D, pairwise infection (contagious3.py:230-248), as a FRONTIER PUSH: only today's infectious agents do
//          work.  For infectious i the transposed interaction_prob row lists every owned u with ip[u,i] != 0; a pair that
//          shares a building today draws one Bernoulli; a success does atomicMax(mailbox[u], i) -- the reference's "last
//          writer wins" is the LARGEST infecting row index (:248), an order-independent max -- and raises u's attention
//          byte.  Two mailboxes per agent: the infector if u moves freely, and the infector if u is isolated at home
//          (status 3: only its home slot counts, :184), so the push does not have to know u's status.
//   agent  everything that is per agent, one streaming pass: B hospitalisation, C mortality, applying D, the ordered
//          rules of E, F household quarantine, G the integer statistics -- and tomorrow's phase A (snapshot byte and
//          the households that WILL see a diagnosis tomorrow, a pure function of an agent's own state and its keyed
//          admission draw).
//
// Because the push of day d+1 only needs the infectious agents' OWN new state (their snapshot byte) and static data,
// a thread block pushes the rows of its own newly infectious agents right after its agent pass of day d: a day of the
// fused kernel (k_days) is  [agent(d) ; push(d+1)]  per block, then ONE grid barrier.  Stand-alone launches (k_push /
// k_agent: injected draws, sharded NCCL mode, first day of a window) find the frontier by scanning the snapshot bytes.
//
// All of it is HBM/L2-bound integer and byte work; there is no dense contraction, hence no tensor-core path.
#pragma once
#include <cstddef>

// Software pipelining of the two gather loops (developer switches for ablation).  Measured on B200 (profiles/r02d_ab.txt):
// requesting the next settle round's records a round early is neutral, requesting the next 64 edge records a step early
// costs 4 % (register pressure under the 64-register cap of 4 CTAs per SM), both together 23 % -- so both are off.
#ifndef DYS_PIPE_SETTLE
#define DYS_PIPE_SETTLE 0
#endif
#ifndef DYS_PIPE_PUSH
#define DYS_PIPE_PUSH 0
#endif
// The two big per-item bodies (one settled agent, one group of pushed rows), or the two block-wide phases, as real
// functions instead of inlined code (developer switches).  Measured on B200 (profiles/r02f_ab.txt, r02g_ab.txt): every
// non-inlined variant is 8-10 % SLOWER than inlining everything, and an L2 prefetch of the records at sweep time buys
// nothing -- all off.
#ifndef DYS_NOINLINE_PUSH
#define DYS_NOINLINE_PUSH 0
#endif
#ifndef DYS_NOINLINE_SETTLE
#define DYS_NOINLINE_SETTLE 0
#endif
// ... or, coarser: the two block-wide PHASES (settle the work list, push the frontier) as real functions, called once per
// flush, so that each gets its own register allocation and the sweep loop around them carries no state of theirs
#ifndef DYS_BLOCK_FNS
#define DYS_BLOCK_FNS 0
#endif
#ifndef DYS_PREFETCH_REC
#define DYS_PREFETCH_REC 0
#endif
#ifndef DYS_BOUNDARY_FIRST
#define DYS_BOUNDARY_FIRST 0
#endif
#ifndef DYS_ROW_REDUX
#define DYS_ROW_REDUX 1
#endif
#if DYS_BLOCK_FNS
#undef DYS_NOINLINE_PUSH
#undef DYS_NOINLINE_SETTLE
#define DYS_NOINLINE_PUSH 0
#define DYS_NOINLINE_SETTLE 0
#define DYS_BLOCK_FN __device__ __noinline__
#else
#define DYS_BLOCK_FN __device__ __forceinline__
#endif
#if DYS_NOINLINE_PUSH
#define DYS_PUSH_FN __device__ __noinline__
#else
#define DYS_PUSH_FN __device__ __forceinline__
#endif
#if DYS_NOINLINE_SETTLE
#define DYS_SETTLE_FN __device__ __noinline__
#else
#define DYS_SETTLE_FN __device__ __forceinline__
#endif

#include "dys_common.cuh"

namespace dys {

// Will agent a (day-start status st, infection day tinf) be among `new_diagnosed_agents` (contagious3.py:261) on
// `day`?  Yes iff it is infectious, day - tinf == diagnosis (rule :254 fires) and phase B does not hospitalise it first.
__device__ __forceinline__ bool diagnosed_on(const DevCtx& c, const double* u_adm, int64_t a, uint32_t st, int tinf, int day,
                                             const double p_adm)
{
    if (!in_set(M_INF, st) || day - tinf != c.diagnosis) return false;
    if (tinf >= c.adm_min && tinf <= c.adm_max) {
        const double u = draw(u_adm, c.lo + a, c.seed, day, STREAM_ADMISSION, (uint32_t)(c.lo + a), 0);
        if (p_adm > u) return false;
    }
    return true;
}

__device__ __forceinline__ void att_or(uint8_t* att, int64_t a, uint32_t bit)
{
    atomicOr(reinterpret_cast<unsigned int*>(att + (a & ~(int64_t)3)), bit << (8 * (int)(a & 3)));
}

// agent a is diagnosed on the day the three targets belong to: every member of its household gets the attention bit
// (contagious3.py:263-265), the quirk households of :267 are flagged, and the day is marked as having a diagnosis
__device__ __forceinline__ void raise_household(const DevCtx& c, int64_t a, int hh_slot, int hh_size, uint8_t* att, uint8_t* qflag,
                                                int32_t* df)
{
    DYS_ASSERT(c, hh_slot >= 0 && hh_size >= 0 && (int64_t)hh_slot + hh_size <= c.n_loc, 6);
    for (int t = 0; t < hh_size; ++t) att_or(att, __ldg(c.hh_mem + hh_slot + t), ATT_HH);
    if (c.trig_ptr)
        for (int64_t t = c.trig_ptr[a]; t < c.trig_ptr[a + 1]; ++t) {
            DYS_ASSERT(c, c.trig_hh[t] >= 0 && c.trig_hh[t] < c.H, 7);
            qflag[c.trig_hh[t]] = 1;                                                             // same-value byte stores
        }
    df[DF_ANY_ND] = 1;
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// peer mode: tomorrow's snapshot bytes of the owned agents that lie in a peer's halo are stored straight into that peer's
// buffer (P2P stores over NVLink), every byte exactly once per day: the sweep stores the bytes of the agents it does not
// list (`listed` = 4-bit mask of the word's agents that go to the settle phase), the settle phase the byte of each agent
// it handles -- so no two stores to one remote byte ever have to be ordered.
__device__ __forceinline__ void remote_ib_word(const DevCtx& c, const int64_t ga0, const uint32_t listed)
{
    for (int p = 1; p < c.n_wr; ++p)
        if (ga0 >= c.pub_lo[p] && ga0 < c.pub_lo[p] + c.pub_n[p]) {
            DYS_ASSERT(c, ga0 >= 0 && ga0 + 4 <= c.ib_stride, 5);
            if (!listed) *reinterpret_cast<uint32_t*>(c.ib_wr[p] + ga0) = 0u;
            else
                for (int j = 0; j < 4; ++j)
                    if (!((listed >> j) & 1u)) c.ib_wr[p][ga0 + j] = 0;
        }
}

__device__ __forceinline__ void remote_ib_byte(const DevCtx& c, const int64_t ga, const uint8_t ibn)
{
    for (int p = 1; p < c.n_wr; ++p)
        if (ga >= c.pub_lo[p] && ga < c.pub_lo[p] + c.pub_n[p]) c.ib_wr[p][ga] = ibn;
}

// Spinning loops (the grid barrier of k_days, the waits for peer GPUs) give
// up when the abort word is set (by the host, or by a peer wait that timed out) or after spin_timeout_ns, so that a
// faulted CTA or a dead peer process ends the launch instead of hanging the GPU.  Returns true when the launch must end.
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <typename Cond>
__device__ __forceinline__ bool spin_until(const DevCtx& c, Cond done)
{
    unsigned spins = 0;
    unsigned long long t0 = 0;
    while (!done()) {
        __nanosleep(32);
        if ((++spins & 255u) == 0u) {
            // (the abort word is DEVICE memory: polling the host's mapped copy from every spinning block cost a PCIe round
            // trip per poll -- 1.4 ms per simulated day on the 8-GPU box, profiles/r02q)
            if (c.abort_word && *c.abort_word) return true;
            const unsigned long long t = global_ns();
            if (!t0) t0 = t;
            else if (c.spin_timeout_ns && t - t0 > c.spin_timeout_ns) {
                if (c.abort_word) { *c.abort_word = 2; *c.abort_host = 2; }      // timed out: everybody else gives up too
                return true;
            }
        }
    }
    return false;
}

// ---------------------------------------------------------------------------------------------------------------
// Shared memory of a block (256 threads = 8 warps).  The frontier outlives the agent pass; plan_cut stages 8192 piece
// costs in front + list.
// ---------------------------------------------------------------------------------------------------------------
constexpr int BLOCK = 256, WARPS = BLOCK / 32;
constexpr int FCAP = 4096;                         // scan-push frontier entries: index relative to the block's base | snapshot byte << 24
constexpr int LTOT = 4096;                         // (with `front`: the staging area of plan_cut)
constexpr uint32_t REL_MASK = 0x00FFFFFFu;

// where a block's queued agents live: halo-relative index = base + rel
struct FrontMap {
    int64_t base;
    __device__ __forceinline__ int64_t at(uint32_t rel) const { return base + rel; }
};

struct PushWarpScratch {
    int off[32];                                   // first flat index of each (non-empty) row of the group
    uint32_t fe[32];                               // its frontier entry
    int64_t t0[32];                                // where its transposed row starts
};

enum { EV_ADM = 0, EV_DEATH, EV_DAILY_INF, EV_DAILY_QUAR, EV_NEW_DIAG, EV_EVENTS, EV_CHANGED, EV_LEN };

struct BlockScratch {
    uint32_t front[FCAP];                          // (front + list: the staging area of plan_cut)
    uint32_t list[LTOT];                           // index relative to the block's base | status << 24 | attention << 28
    PushWarpScratch pw[WARPS];
    double pdf[64];
    long long b_lo[2];                             // the batch being swept and the one after it (first piece, pieces)
    int b_n[2];
    unsigned int n_front, n_list;
    unsigned int push_acc[4];                      // the block's push counters: entries scanned, candidates, Bernoulli draws, successes
    unsigned int pub_cnt[8];                       // peer mode: pieces of this call whose bytes went to write target p
    unsigned int hist[HIST_LEN], ev[EV_LEN], cnt[10], acc[8];
    bool last;
};

// ---------------------------------------------------------------------------------------------------------------
// push_group: ONE WARP pushes the transposed rows of <= 32 queued infectious agents (ent[0 .. nb), halo-relative index =
// fmap.at(entry & REL_MASK), snapshot byte = entry >> 24).  Empty rows are dropped, the others are flattened (warp prefix
// sum of the row lengths) so that every lane works on consecutive 16-byte records however long the rows are, two per
// lane in flight.  The row of a flat index is found with one ballot and one REDUX per 32 entries (rows that end at or
// before the chunk + a bit mask of the row ends inside it), not a search per entry.
//   CHECK = true   the status column is final (stand-alone launch, or after a grid barrier): u must be susceptible, and
//                  the hit goes to the mailbox that matches u's isolation -- the counters equal the reference's
//                  candidate / exposed / success counts.
//   CHECK = false  the push runs while other blocks still settle day d: nothing about u is read.  A hit is recorded
//                  for both isolation cases; the agent pass of u picks the one that applies (and ignores it if u is
//                  not susceptible).
// ---------------------------------------------------------------------------------------------------------------
// one lane's row of a group: frontier entry, start and length of its transposed row
struct RowRef {
    uint32_t fe;
    int len;
    int64_t t0;
};

__device__ __forceinline__ RowRef load_row(const DevCtx& c, const uint32_t* ent, const int nb, const FrontMap fmap)
{
    RowRef r{0u, 0, 0};
    const int lane = threadIdx.x & 31;
    if (lane < nb) {
        r.fe = ent[lane];
        const int64_t il = fmap.at(r.fe & REL_MASK);
        r.t0 = __ldg(c.tptr + il);
        r.len = (int)(__ldg(c.tptr + il + 1) - r.t0);
    }
    return r;
}

template <bool CHECK>
DYS_PUSH_FN uint4 push_group(const DevCtx& c, const int day, PushWarpScratch& w, const double* __restrict__ s_pdf,
                             const RowRef row0, const FrontMap fmap)
{
    unsigned int n_cand = 0u, n_expo = 0u, n_hit = 0u;      // candidates, Bernoulli draws, successes of the whole group
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const bool any_closed = c.n_closed != 0;
    const uint4* edges = reinterpret_cast<const uint4*>(c.tedge);
    const int64_t t0 = row0.t0;
    const int len = row0.len;
    const uint32_t fe0 = row0.fe;
    const unsigned nz = __ballot_sync(0xffffffffu, len > 0);
    const int nr = __popc(nz);
    if (nr == 0) return make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();                                    // (the previous group's readers are done with the scratch)
    if (len > 0) {
        const int slot = __popc(nz & lt_mask);
        w.t0[slot] = t0;
        w.fe[slot] = fe0;
        w.off[slot] = len;
    }
    __syncwarp();
    const int mylen = lane < nr ? w.off[lane] : 0;
    int end = mylen;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, end, o);
        if (lane >= o) end += v;
    }
    const int total = __shfl_sync(0xffffffffu, end, 31);
    __syncwarp();
    w.off[lane] = end - mylen;
    __syncwarp();
    // two entries per lane and step: every entry of the transposed network can be exposed (the others were left out at
    // load time), so nearly every lane draws its Bernoulli and the Philox rounds run on full warps
    // (DYS_PIPE_PUSH: the records of step f0 + 64 are requested before the pairs of step f0 are evaluated)
#if DYS_PIPE_PUSH
    uint32_t fe_n[2];
    uint4 rec_n[2];
#endif
    auto fetch = [&](const int f0, uint32_t (&fe)[2], uint4 (&rec)[2]) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int base = f0 + 32 * k, f = base + lane;
#if DYS_ROW_REDUX
            // row of f = #rows that end at or before f (row ends are strictly increasing: empty rows were dropped)
            const int before = __popc(__ballot_sync(0xffffffffu, lane < nr && end <= base));
            const unsigned ends = __reduce_or_sync(0xffffffffu, (lane < nr && end > base && end <= base + 31) ? 1u << (end - base) : 0u);
            const int row = before + __popc(ends & ((2u << lane) - 1u));
#else
            int lo = 0, hi = 32;                       // off[lo] <= f < off[hi]; the last such lo is the row
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const int mid = (lo + hi) >> 1;
                if (mid < nr && w.off[mid] <= f) lo = mid; else hi = mid;
            }
            const int row = lo;
#endif
            rec[k] = make_uint4(0u, 0u, 0u, 0u);       // class 0: a padding lane never passes the exposure test
            fe[k] = 0u;
            if (f < total) {
                DYS_ASSERT(c, row >= 0 && row < nr && f >= w.off[row], 1);
                DYS_ASSERT(c, w.t0[row] + (f - w.off[row]) >= 0 && w.t0[row] + (f - w.off[row]) < __ldg(c.tptr + c.rt_n), 2);
                rec[k] = __ldg(edges + (w.t0[row] + (f - w.off[row])));
                fe[k] = w.fe[row];
            }
        }
    };
#if DYS_PIPE_PUSH
    fetch(0, fe_n, rec_n);
#endif
    for (int f0 = 0; f0 < total; f0 += 64) {
#if DYS_PIPE_PUSH
        uint32_t fe[2] = {fe_n[0], fe_n[1]};
        uint4 rec[2] = {rec_n[0], rec_n[1]};
        if (f0 + 64 < total) fetch(f0 + 64, fe_n, rec_n);
#else
        uint32_t fe[2];
        uint4 rec[2];
        fetch(f0, fe, rec);
#endif
        uint32_t st_u[2];
        if (CHECK) {
#pragma unroll
            for (int k = 0; k < 2; ++k)
                st_u[k] = (f0 + 32 * k + lane) < total ? (uint32_t)c.status[rec[k].x & COL_MASK] : (uint32_t)ST_PAD;   // day-start status
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            // the exposure test: static class x today's isolation
            const uint32_t cls = rec[k].x >> CLS_SHIFT, ib_i = fe[k] >> 24;
            const bool iso_i = (ib_i & IB_ISO) != 0;
            const bool x_iso = (cls & CLS_HH) | ((cls & CLS_HO) && !iso_i);                      // u isolated at home
            const bool x_free = x_iso | ((cls & CLS_OH) != 0) | ((cls & CLS_OO) && !iso_i);      // u moves freely
            // (the counters are warp-uniform -- a ballot and a popcount per 32 pairs on the uniform datapath -- instead of four
            // per-lane registers carried through the loop: the day kernel lives under a 64-register cap)
            bool ok_free, ok_iso, iso_u = false, cand;
            if (CHECK) {
                const bool sus = in_set(M_SUS, st_u[k]);
                iso_u = (st_u[k] == ST_QH);
                cand = sus;
                ok_free = sus && !iso_u && x_free;
                ok_iso = sus && iso_u && x_iso;
            } else {
                cand = x_free;
                ok_free = x_free;
                ok_iso = x_iso;
            }
            n_cand += __popc(__ballot_sync(0xffffffffu, cand));
            bool expo = ok_free || ok_iso, hit = false;
            if (expo) {
                const uint32_t ul = rec[k].x & COL_MASK;
                const int64_t ug = c.lo + ul, ig = c.rt_lo + fmap.at(fe[k] & REL_MASK);
                DYS_ASSERT(c, (int64_t)ul < c.n_loc && ig >= c.rt_lo && ig < c.rt_lo + c.rt_n, 3);
                if (any_closed && (rec[k].y & EDGE_ONE_BLD)) {
                    // closed buildings, and the pair shares ONE building: the class test above already is the exposure test
                    // for that building -- it only has to be open (one byte from a table that lives in L2)
                    const bool open_b = __ldcg(c.open + (rec[k].y & ~EDGE_ONE_BLD)) != 0;
                    ok_free = ok_free && open_b;
                    ok_iso = ok_iso && open_b;
                    expo = ok_free || ok_iso;
                } else if (any_closed && (rec[k].y & EDGE_MASKS)) {      // ... several buildings: which slots of u meet i
                    if (CHECK) {
                        const bool ok = exposed_masks(c.routine, c.open, c.V, ug - c.rt_lo, rec[k].y, iso_u, iso_i);
                        ok_free = ok_free && ok;
                        ok_iso = ok_iso && ok;
                    } else {
                        ok_free = exposed_masks(c.routine, c.open, c.V, ug - c.rt_lo, rec[k].y, false, iso_i);
                        ok_iso = ok_iso && exposed_masks(c.routine, c.open, c.V, ug - c.rt_lo, rec[k].y, true, iso_i);
                    }
                    expo = ok_free || ok_iso;
                } else if (any_closed) {      // (V > 10: no room for the masks) re-check the routines
                    if (CHECK) {
                        const bool ok = exposed_slow(c.routine, c.open, c.V, ug - c.rt_lo, ig - c.rt_lo, iso_u, iso_i);
                        ok_free = ok_free && ok;
                        ok_iso = ok_iso && ok;
                    } else {
                        ok_free = exposed_slow(c.routine, c.open, c.V, ug - c.rt_lo, ig - c.rt_lo, false, iso_i);
                        ok_iso = ok_iso && exposed_slow(c.routine, c.open, c.V, ug - c.rt_lo, ig - c.rt_lo, true, iso_i);
                    }
                    expo = ok_free || ok_iso;
                }
                if (expo) {
                    // contagious3.py:232-238: ((ip[u,i] * risk[u]) * pdf(days since i's infection)) * norm_factor, each product
                    // rounded separately (no add follows, so no FMA can form); the first product comes with the record
                    double p = __dmul_rn(__hiloint2double((int)rec[k].w, (int)rec[k].z), s_pdf[ib_i & IB_DAYS]);
                    p = __dmul_rn(p, c.norm_factor);
                    if (CHECK && c.pair_prob) c.pair_prob[(size_t)ug * c.n_glob + ig] = p;
                    const double u = draw(c.u_pair, ug * c.n_glob + ig, c.seed, day, STREAM_INFECTION, (uint32_t)ug, (uint32_t)ig);
                    if (p > u) {                                                       // :240, :248
                        hit = true;
                        if (ok_free) atomicMax(&c.mb_any[ul], (int)ig);
                        if (ok_iso) atomicMax(&c.mb_iso[ul], (int)ig);
                        att_or(c.att, ul, ATT_HIT);
                    }
                }
            }
            n_expo += __popc(__ballot_sync(0xffffffffu, expo));
            n_hit += __popc(__ballot_sync(0xffffffffu, hit));
        }
    }
    return make_uint4((unsigned)total, n_cand, n_expo, n_hit);      // (the same four numbers in every lane)
}

// the block-wide frontier (sm.front[0 .. n_front), halo-relative index = fmap_base + rel), dealt to the warps in groups of
// <= 32 rows; the counters go to sm.push_acc
template <bool CHECK>
DYS_BLOCK_FN void push_frontier(const DevCtx& c, const int day, BlockScratch& sm, const int64_t fmap_base)
{
    const FrontMap fmap{fmap_base};
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = (int)sm.n_front;
    int rpg = (n + WARPS - 1) / WARPS;
    rpg = rpg < 1 ? 1 : (rpg > 32 ? 32 : rpg);
    const int n_groups = (n + rpg - 1) / rpg;
    auto rows_of = [&](const int g) {
        const int q0 = g * rpg;
        const int nb = (n - q0) < rpg ? (n - q0) : rpg;
        return load_row(c, sm.front + q0, nb, fmap);
    };
    unsigned int cnt[4] = {0u, 0u, 0u, 0u};
#if DYS_PIPE_PUSH
    RowRef next = wid < n_groups ? rows_of(wid) : RowRef{0u, 0, 0};
#endif
    for (int g = wid; g < n_groups; g += WARPS) {
#if DYS_PIPE_PUSH
        const RowRef cur = next;
        if (g + WARPS < n_groups) next = rows_of(g + WARPS);      // the next group's row extents, in flight under this group
#else
        const RowRef cur = rows_of(g);
#endif
        const uint4 d = push_group<CHECK>(c, day, sm.pw[wid], sm.pdf, cur, fmap);
        cnt[0] += d.x; cnt[1] += d.y; cnt[2] += d.z; cnt[3] += d.w;
    }
    if (lane == 0)      // (push_group's counts are per group, the same in every lane)
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (cnt[k]) atomicAdd(&sm.push_acc[k], cnt[k]);
}

// what a push of day X clears before the agent pass of day X fills it / raises into it (grid-strided, any block)
// (c = the view of replica c.rep; r = another replica of the same day, addressed relative to it)
__device__ __forceinline__ void push_clear_duties(const DevCtx& c, const int r = -1)
{
    const int dr = r < 0 ? 0 : r - c.rep;
    const ptrdiff_t o = (ptrdiff_t)c.out_stride * dr;
    if (c.sa_hist) grid_clear(reinterpret_cast<int32_t*>(reinterpret_cast<unsigned char*>(c.sa_hist) + o), c.S_out * HIST_LEN);
    if (c.bcount) grid_clear(reinterpret_cast<int32_t*>(reinterpret_cast<unsigned char*>(c.bcount) + o), c.B_out);
    grid_clear(c.io_mat, c.io_mat ? c.S * c.S : 0);
    if (c.trig_ptr) grid_clear16(c.qflag_next + (ptrdiff_t)(c.H + 16) * dr, c.H);
    if (blockIdx.x == 0 && threadIdx.x < DF_LEN) c.df_next[DF_LEN * dr + (int)threadIdx.x] = 0;
}

// the block's push counters (sm.push_acc) into the day's spread partial sums; block-synchronising
__device__ __forceinline__ void publish_push_counters(const DevCtx& c, BlockScratch& sm)
{
    __syncthreads();
    if (threadIdx.x < 4 && sm.push_acc[threadIdx.x]) {
        const int idx = threadIdx.x == 0 ? ST_EDGES_SCANNED : threadIdx.x == 1 ? ST_CANDIDATES : threadIdx.x == 2 ? ST_EXPOSED : ST_HITS;
        atomicAdd(&c.part[(size_t)idx * N_SLOTS + (blockIdx.x & (N_SLOTS - 1))], (unsigned long long)sm.push_acc[threadIdx.x]);
        sm.push_acc[threadIdx.x] = 0u;
    }
    __syncthreads();
}

// the block's contiguous share [lo, hi) of n items, in granules of 16
__device__ __forceinline__ void block_share(const int64_t n, const int n_blocks, const int wb, int64_t& lo, int64_t& hi)
{
    if (wb < 0) { lo = hi = 0; return; }
    int64_t span = (n + n_blocks - 1) / n_blocks;
    span = (span + 15) & ~(int64_t)15;
    lo = (int64_t)wb * span;
    hi = lo + span;
    if (lo > n) lo = n;
    if (hi > n) hi = n;
}

// ---------------------------------------------------------------------------------------------------------------
// push_scan_body: the push of `day` with the frontier found by scanning the snapshot bytes (c.ib_rd) of the halo
// agents.  Every block scans a contiguous share, 1024 bytes a step (the next step's word already in flight).
// ---------------------------------------------------------------------------------------------------------------
// wb = the block's index among the n_blocks working blocks (-1: a block without a share)
// skip_own (peer mode inside a window): the owned agents' rows were already pushed by their blocks after yesterday's
// agent pass; only the halo agents' bytes, which arrive from the peers, are scanned
// returns true when a wait for a peer GPU was abandoned (abort word / timeout): the launch must end
// HALO_ONLY (peer mode inside a window): the rows of the owned agents were pushed by their blocks after the agent pass; only
// the halo agents' bytes, which the peers stored into this GPU's buffer, are scanned -- while other blocks may still be
// settling the day before, hence CHECK = false and no clearing duties
template <bool HALO_ONLY, bool PEERS>
__device__ __forceinline__ bool push_scan_body(const DevCtx& c, const int day, BlockScratch& sm, const int n_blocks, const int wb)
{
    constexpr bool CHECK = !HALO_ONLY;
    __shared__ int s_gave_up;
    const int tid = threadIdx.x, lane = tid & 31;
    if (!HALO_ONLY) push_clear_duties(c);
    if (tid < 64) sm.pdf[tid] = c.pdf[tid];
    if (tid == 0) sm.n_front = 0;
    if (tid < 4) sm.push_acc[tid] = 0u;
    __syncthreads();

    // halo-relative agents [r0, r1): r0 is a multiple of 1024
    auto scan = [&](const int64_t r0, const int64_t r1) {
        int64_t lo, hi;
        block_share(r1 - r0, n_blocks, wb, lo, hi);
        lo += r0; hi += r0;
        int64_t base = lo;
        const uint32_t* ib4 = reinterpret_cast<const uint32_t*>(c.ib_rd + c.rt_lo);   // rt_lo is a multiple of 1024
        uint32_t v = 0;
        // (__ldcg: the bytes were written by other SMs or other GPUs since this SM last read the buffer; L1 is not coherent)
        if (lo + tid * 4 < hi) v = __ldcg(ib4 + (lo >> 2) + tid);
        for (int64_t s0 = lo; s0 < hi; s0 += BLOCK * 4) {
            uint32_t vn = 0;
            if (s0 + BLOCK * 4 + tid * 4 < hi) vn = __ldcg(ib4 + ((s0 + BLOCK * 4) >> 2) + tid);   // next step's bytes, in flight
            const int64_t il0 = s0 + tid * 4;
            uint32_t vm = v & 0x80808080u;                          // bytes past the end of the share are not ours
            if (il0 + 4 > hi) vm &= il0 >= hi ? 0u : (0xFFFFFFFFu >> (8 * (int)(il0 + 4 - hi)));
            const int mine = __popc(vm);
            const unsigned any = __ballot_sync(0xffffffffu, mine != 0);
            if (any) {
                int incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                int pos = 0;
                if (lane == 31) pos = (int)atomicAdd(&sm.n_front, (unsigned)incl);
                pos = __shfl_sync(0xffffffffu, pos, 31) + incl - mine;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const uint32_t byte = (v >> (8 * b)) & 0xFFu;
                    if ((vm >> (8 * b)) & 0x80u) {
                        DYS_ASSERT(c, pos < FCAP && il0 + b - base <= (int64_t)REL_MASK, 4);
                        sm.front[pos++] = (uint32_t)(il0 + b - base) | (byte << 24);
                        prefetch_l2(c.tptr + il0 + b);          // the row extent is needed when the frontier is pushed
                    }
                }
            }
            v = vn;
            __syncthreads();
            const int64_t s1 = s0 + BLOCK * 4;
            const unsigned nf = sm.n_front;        // read by everyone before the next step appends (barrier below)
            if (nf > FCAP - BLOCK * 4 || s1 - base > (int64_t)REL_MASK - BLOCK * 4 || s1 >= hi) {
                if (nf) push_frontier<CHECK>(c, day, sm, base);
                __syncthreads();
                if (tid == 0) sm.n_front = 0;
                base = s1;
            }
            __syncthreads();
        }
    };
    if (PEERS && c.n_peers > 1) {
        // peer mode: the owned agents' bytes are local and final, so their rows are pushed first; the halo agents' bytes
        // were stored into this GPU's buffer by their owners, who counted what they delivered in arrive[r] (pieces + one
        // token per day): the bytes of this day are complete once epoch x expect_in[r] units have arrived.  The exchange
        // latency hides behind the local part of the push.  (Persistent grid: all CTAs resident, spinning is safe.)
        const int64_t own0 = c.lo - c.rt_lo, own1 = own0 + c.n_pad < c.rt_n ? own0 + c.n_pad : c.rt_n;
        if (!HALO_ONLY) scan(own0, own1);
        if (tid == 0) s_gave_up = 0;
        __syncthreads();
        if (tid < c.n_peers && c.expect_in[tid]) {
            const volatile unsigned long long* f = c.arrive + tid;
            const unsigned long long need = c.epoch * (unsigned long long)c.expect_in[tid];
            if (spin_until(c, [&]() { return *f >= need; })) s_gave_up = 1;
            __threadfence_system();
        }
        __syncthreads();
        if (s_gave_up) return true;
        if (own0 > 0) scan(0, own0);
        if (own1 < c.rt_n) scan(own1, c.rt_n);
    } else {
        scan(0, c.rt_n);
    }
    publish_push_counters(c, sm);
    return false;
}

__global__ void __launch_bounds__(BLOCK) k_push(const DevCtx c, const int day)
{
    __shared__ BlockScratch sm;
    __shared__ DevCtx s_c;
    if (threadIdx.x == 0) s_c = c;
    __syncthreads();
    push_scan_body<false, true>(s_c, day, sm, (int)gridDim.x, (int)blockIdx.x);
}

// ---------------------------------------------------------------------------------------------------------------
// agent_body.  A block owns a contiguous share of the agents and walks it in batches of 2048, in two phases.
//   sweep  : 4 agents per thread and round: one 32-bit load of the status bytes, one of the attention bytes.  An agent
//            to whom nothing can happen today -- susceptible or long recovered, attention byte 0 -- only feeds the
//            packed status histogram.  Every other agent is compacted (ballot + prefix popcount) into the work list.
//   settle : the work list is processed densely, one agent per thread: B, C, applying D, the ordered rules of E, F, G,
//            tomorrow's snapshot byte and household flags; agents that are infectious tomorrow join the block's frontier.
// Divergent per-agent logic therefore costs instructions in proportion to the agents that need it, not to the
// population.  Compartment counts are reduced as packed 8-bit histograms (REDUX); rare events go to shared-memory
// counters; one block folds the spread partial sums into the day's stats block.
// ---------------------------------------------------------------------------------------------------------------
// packed 8-bit histogram field of a status code: QI S I QH | D H R X
__device__ __forceinline__ uint32_t hist_idx(uint32_t st) { return (st == ST_QI) ? 0u : (st >> 1); }
// internal partial-sum rows (folded into the public stats by fold_stats)
constexpr int RAW_END = 53;      // [53..60] end-of-day status counts in hist_idx order
constexpr int RAW_SUS0 = 61, RAW_INF0 = 62;
constexpr int ST_COUNT0 = 20;    // public: stats[20..27] = end-of-day agents in S, I, QH, QI, D, H, R, X

// what one settled agent contributes to the day's counters (accumulated per warp, see agent_body)
enum : uint32_t { EVB_ADM = 1u, EVB_DEATH = 2u, EVB_DAILY_INF = 4u, EVB_DAILY_QUAR = 8u, EVB_NEW_DIAG = 16u, EVB_CHANGED = 32u,
                  EVB_SUS0 = 64u, EVB_INF0 = 128u };
struct Settled {
    uint32_t x;          // end-of-day status
    uint32_t ev;         // EVB_* flags
    int k;               // days-since-infection histogram bin, or -1
    uint32_t ibn;        // tomorrow's snapshot byte
    int infector;        // >= 0: infected today by this agent (global index)
};

// what the settle phase fetches for one agent, all of it requested before any of it is used
struct Loaded {
    int4 rv;             // first half of the agent record: days in state, household size, area, home building
    int mb;              // the infector the push left for this agent, or -1
};

__device__ __forceinline__ Loaded load_agent(const DevCtx& c, const int64_t a, const uint32_t st, const uint32_t att)
{
    Loaded L;
    // (the second half -- household, member list, p_adm -- is a dependent load on the few days of an agent's life when it
    // can be admitted or diagnosed: fetching both halves for everybody measured 4 % slower)
    L.rv = reinterpret_cast<const int4*>(c.rec + a)[0];
    L.mb = -1;
    if (att & ATT_HIT) {
        // the mailbox that matches this agent's isolation; both are reset for the day after tomorrow
        const int m_free = c.mb_any[a], m_iso = c.mb_iso[a];
        L.mb = (st == ST_QH) ? m_iso : m_free;
        c.mb_any[a] = -1;
        c.mb_iso[a] = -1;
    }
    return L;
}

// one non-idle agent, start to end of day (contagious3.py:205-270 for this agent, then its share of :273-366)
template <bool PEERS>
DYS_SETTLE_FN Settled settle_agent(const DevCtx& c, const int day, const int64_t a, const bool any_nd,
                                   const uint32_t st, const uint32_t att, const Loaded L)
{
    bool infected_today = false;
    int infector = -1;
    uint32_t evb = 0;
    const int4 rv = L.rv;
    int16_t* const rdays = reinterpret_cast<int16_t*>(c.rec + a);       // [0] t_inf, [1] t_q, [2] t_adm
    int ti = (int16_t)(rv.x & 0xFFFF);
    int tq = rv.x >> 16;
    const int tadm = (int16_t)(rv.y & 0xFFFF), hh_size = rv.y >> 16;
    const int sa_a = rv.z, home_a = rv.w;
    auto second_half = [&]() { return reinterpret_cast<const int4*>(c.rec + a)[1]; };
    const bool quirk = any_nd && (c.trig_ptr || c.hh_always);
    const int h = quirk ? second_half().x : 0;
    const int mb = L.mb;
    const bool f_hh = (att & ATT_HH) != 0;                                                  // :263
    bool f_q = false;
    if (quirk) f_q = (c.trig_ptr && c.qflag[h] != 0) || (c.hh_always && __ldg(c.hh_always + h));   // qflag is written during k_days: no __ldg
    uint32_t x = st;
    if (!in_set(M_EVER, st)) ti = 0;
    bool w_ti = false, w_tq = false;
    infected_today = false;
    // ---- B, C (contagious3.py:205-227) on the day-start status; D's result from the mailbox ----
    if (in_set(M_INF, st)) {
        evb |= EVB_INF0;
        if (ti >= c.adm_min && ti <= c.adm_max) {                            // :205 window on the ABSOLUTE infection day
            const double u = draw(c.u_adm, c.lo + a, c.seed, day, STREAM_ADMISSION, (uint32_t)(c.lo + a), 0);
            const int4 rw = second_half();
            if (__hiloint2double(rw.w, rw.z) > u) { x = ST_H; rdays[2] = (int16_t)day; evb |= EVB_ADM; }   // :213-216
        }
    } else if (st == ST_H) {
        if (tadm >= c.death_min) {                                           // :219
            const double u = draw(c.u_death, c.lo + a, c.seed, day, STREAM_DEATH, (uint32_t)(c.lo + a), 0);
            if (__ldg(c.p_death + a) > u) { x = ST_X; rdays[2] = 0; evb |= EVB_DEATH; }            // :225-227, :259
        }
    } else if (in_set(M_SUS, st)) {
        evb |= EVB_SUS0;
        infector = mb;
        if (infector >= 0) {                                                 // :241-248: the push found an infector
            x = (st == ST_S) ? ST_I : ST_QI;
            ti = day; w_ti = true;
            infected_today = true;
            c.src[a] = infector;
        }
    }
    // ---- E: the ordered rules of contagious3.py:250-259 ----
    const int n_inf = day - ti, n_q = day - tq;
    if (x == ST_QH && n_q == c.quarantine) x = ST_S;                                        // :250
    if (x == ST_QI && n_q == c.quarantine) x = ST_I;                                        // :252
    if (x == ST_I && n_inf == c.diagnosis) { tq = day; w_tq = true; }                       // :253
    if ((x == ST_I || x == ST_QI) && n_inf == c.diagnosis) x = ST_D;                        // :254
    if (x == ST_D && n_inf == c.recover) x = ST_R;                                          // :255
    if (x == ST_H && n_inf == c.hospital_recover) { x = ST_R; rdays[2] = 0; }               // :256, :259
    if (x == ST_D && n_inf == c.diagnosis) evb |= EVB_NEW_DIAG;                             // :261
    // ---- F (contagious3.py:261-270): the households diagnosed today were flagged yesterday ----
    if (any_nd && (x == ST_S || x == ST_I)) {
        bool hit = f_hh;                                                                    // :263
        if (x == ST_I && !hit) hit = f_q;                                                   // :267 whole-row isin
        if (hit && c.compliance < 1.0)
            hit = keyed_uniform(c.seed, day, STREAM_COMPLIANCE, (uint32_t)(c.lo + a), 0) < c.compliance;
        if (hit) { x = (x == ST_S) ? ST_QH : ST_QI; tq = day; w_tq = true; }                // :265-270
    }
    if (w_ti) rdays[0] = (int16_t)ti;
    if (w_tq) rdays[1] = (int16_t)tq;
    if (x != st) { c.status[a] = (uint8_t)x; evb |= EVB_CHANGED; }
    if (in_set(M_ISO, x) && tq == day) evb |= EVB_DAILY_QUAR;
    // ---- G: this agent's share of the integer statistics (contagious3.py:273-366) ----
    int kbin = -1;
    if (in_set(M_EVER, x)) {
        const int k = day - ti;
        if (k >= 0 && k < HIST_LEN) {
            if (k == 0) evb |= EVB_DAILY_INF;
            kbin = k;
            DYS_ASSERT(c, sa_a < c.S_out && home_a >= 0 && home_a < c.B_out, 8);
            if (c.sa_hist && sa_a >= 0) atomicAdd(&c.sa_hist[(size_t)sa_a * HIST_LEN + k], 1);
            if (c.io_mat && infected_today) {
                // contagious3.py:356-366: row = area of today's infected, column = area of the infector
                const int sa_u = sa_a, sa_i = c.rec[infector - c.lo].sa;
                if (sa_u >= 0 && sa_i >= 0) atomicAdd(&c.io_mat[(size_t)sa_u * c.S + sa_i], 1);
            }
        }
        if (c.bcount && in_set(M_INF, x)) atomicAdd(&c.bcount[home_a], 1);
    }
    // ---- tomorrow's phase A: snapshot byte, and the households that will see a diagnosis tomorrow ----
    const uint8_t ibn = infectious_byte(x, day + 1, ti);
    if (ibn && c.write_ib) c.ib_wr[0][c.lo + a] = ibn;      // (the sweep already stored 0 for everybody)
    if (PEERS && c.n_wr > 1) remote_ib_byte(c, c.lo + a, ibn);
    if (in_set(M_INF, x) && day + 1 - ti == c.diagnosis) {
        const int4 rw = second_half();
        if (diagnosed_on(c, nullptr, a, x, ti, day + 1, __hiloint2double(rw.w, rw.z)))
            raise_household(c, a, rw.y, hh_size, c.att_next, c.qflag_next, c.df_next);
    }
    return Settled{x, evb, kbin, ibn, infected_today ? infector : -1};
}

// One block folds the spread partial sums (part[counter][slot]: one warp per counter, two coalesced loads, shuffle tree),
// derives the compartment statistics from the status counts and re-zeroes the partials (this day's ring entry).
__device__ __forceinline__ void fold_stats(unsigned long long* __restrict__ part, int64_t* __restrict__ stats)
{
    __shared__ long long s_fold[N_STATS];
    const int tid = threadIdx.x, lane = tid & 31;
    __threadfence();
    unsigned long long t[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        unsigned long long* p = part + (size_t)((tid >> 5) + 8 * q) * N_SLOTS;
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
        stats[tid] = v;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------
// Where a block's pieces come from.
//   static  (word == null): the pieces [next, end) of the view, in batches of <= 16 (k_agent; a stolen chunk in k_days).
//   dynamic (k_days, the block's own run): `word` = (end << 32 | next) in UNIFIED piece indices ([replica][piece], local =
//           unified - rep_off) lives in global memory.  The owner takes batches from the head with a compare-and-swap; blocks
//           that finished their own run take batches from the same word (work stealing, see k_days): one atomicAdd per
//           batch -- 16 pieces for the owner, 8 for a thief --, issued a whole sweep before its answer is needed.
// ---------------------------------------------------------------------------------------------------------------
struct WorkSrc {
    unsigned long long* word;
    int64_t rep_off;
    int64_t next, end;                   // static mode: local piece indices
};

__device__ __forceinline__ unsigned long long pack_work(const int64_t next, const int64_t end)
{
    return ((unsigned long long)(uint32_t)end << 32) | (unsigned long long)(uint32_t)next;
}

// peer mode, one thread: the pieces [lo, lo + n) of the view were taken by this block -- how many of them lie in each peer's halo
__device__ __forceinline__ void count_pub(const DevCtx& c, BlockScratch& sm, const int64_t lo, const int n)
{
    const int64_t g0 = c.lo + (lo << 7), g1 = g0 + ((int64_t)n << 7);
    for (int p = 1; p < c.n_wr; ++p) {
        const int64_t a = g0 > c.pub_lo[p] ? g0 : c.pub_lo[p];
        const int64_t b = g1 < c.pub_lo[p] + c.pub_n[p] ? g1 : c.pub_lo[p] + c.pub_n[p];
        if (b > a) sm.pub_cnt[p] += (unsigned)((b - a) >> 7);
    }
}

// one thread: take the next batch of <= g pieces.  The atomic is ISSUED by take_issue and its answer only looked at by
// take_result, so that the round trip to the work word hides under whatever the caller does in between.
__device__ __forceinline__ unsigned long long take_issue(WorkSrc& ws, const int g)
{
    if (!ws.word) {
        const int64_t lo = ws.next;
        const int64_t n = ws.end - lo < g ? ws.end - lo : g;
        ws.next = lo + (n > 0 ? n : 0);
        return pack_work(lo, lo + (n > 0 ? n : 0));
    }
    const unsigned long long old = atomicAdd(ws.word, (unsigned long long)g);
    const uint32_t nx = (uint32_t)old, en = (uint32_t)(old >> 32);
    const uint32_t hi = nx + (uint32_t)g < en ? nx + (uint32_t)g : en;
    return nx < en ? pack_work((int64_t)nx - ws.rep_off, (int64_t)hi - ws.rep_off) : 0ull;      // local [lo, hi)
}

__device__ __forceinline__ int64_t take_result(const unsigned long long t, int& n)
{
    const uint32_t lo = (uint32_t)t, hi = (uint32_t)(t >> 32);
    n = hi > lo ? (int)(hi - lo) : 0;
    return (int64_t)lo;
}

// settle_list: the block's work list (sm.list[0 .. n_list)) processed densely, one agent per thread whatever piece it came
// from; agents that are infectious tomorrow are appended to the block's frontier (cn != null)
template <int VARIANT, bool PEERS>
DYS_BLOCK_FN void settle_list(const DevCtx& c, const int day, BlockScratch& sm, const DevCtx* cn, const int64_t a_base, const bool any_nd)
{
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const FrontMap fmap{c.lo - c.rt_lo + a_base};
    const int n_list = (int)sm.n_list;
#if DYS_PIPE_SETTLE
    // the next round's records are requested before this round is settled
    uint32_t le_n = tid < n_list ? sm.list[tid] : 0u;
    Loaded L_n;
    L_n.mb = -1;
    if (tid < n_list) L_n = load_agent(c, a_base + (le_n & REL_MASK), (le_n >> 24) & 0xFu, (le_n >> 28) & 0x3u);
#endif
    for (int k0 = 0; k0 < n_list; k0 += BLOCK) {
        const int k = k0 + tid;
#if !DYS_PIPE_SETTLE
        uint32_t le_n = 0u;
        Loaded L_n;
        L_n.mb = -1;
        if (k < n_list) {
            le_n = sm.list[k];
            L_n = load_agent(c, a_base + (le_n & REL_MASK), (le_n >> 24) & 0xFu, (le_n >> 28) & 0x3u);
        }
#endif
        bool infected_today = false;
        int infector = -1;
        uint32_t ibn = 0;
        uint32_t h0 = 0, h1 = 0, eA = 0, eB = 0;       // this agent's counter contributions, 8-bit fields
        int kbin = -1;
        const bool valid = k < n_list;
        const uint32_t le = le_n;
        const Loaded L = L_n;
        const uint32_t rel = le & REL_MASK;
        const int64_t a = a_base + rel;
        const uint32_t st = valid ? (le >> 24) & 0xFu : (uint32_t)ST_PAD, att = (le >> 28) & 0x3u;
#if DYS_PIPE_SETTLE
        L_n.mb = -1;
        if (k + BLOCK < n_list) {
            le_n = sm.list[k + BLOCK];
            L_n = load_agent(c, a_base + (le_n & REL_MASK), (le_n >> 24) & 0xFu, (le_n >> 28) & 0x3u);
        }
#endif
        // today's infection events: one slot claim per warp that has any (order is arbitrary; hosts sort).  The claim
        // is issued as soon as the mailboxes are in, its answer is only needed when the events are written
        const unsigned evm = __ballot_sync(0xffffffffu, valid && in_set(M_SUS, st) && L.mb >= 0);
        int ev_base = 0;
        if (evm && c.ev && lane == 0) ev_base = atomicAdd(&c.df[DF_N_EVENTS], __popc(evm));
        if (valid) {
            const Settled r = settle_agent<PEERS>(c, day, a, any_nd, st, att, L);
            infector = r.infector;
            infected_today = infector >= 0;
            ibn = r.ibn;
            kbin = r.k;
            const uint32_t hi = hist_idx(r.x);
            h0 = hi < 4 ? 1u << (8 * hi) : 0u;
            h1 = hi < 4 ? 0u : 1u << (8 * (hi - 4));
            eA = ((r.ev & EVB_ADM) ? 1u : 0u) | ((r.ev & EVB_DEATH) ? 1u << 8 : 0u) | ((r.ev & EVB_DAILY_INF) ? 1u << 16 : 0u) |
                 ((r.ev & EVB_DAILY_QUAR) ? 1u << 24 : 0u);
            eB = ((r.ev & EVB_NEW_DIAG) ? 1u : 0u) | ((r.ev & EVB_CHANGED) ? 1u << 8 : 0u) | ((r.ev & EVB_SUS0) ? 1u << 16 : 0u) |
                 ((r.ev & EVB_INF0) ? 1u << 24 : 0u);
        }
        {
            // the warp's counters in four REDUX (at most 32 per 8-bit field), then one shared-memory atomic per
            // non-zero counter from distinct lanes -- not 32 colliding atomics per counter
            h0 = __reduce_add_sync(0xffffffffu, h0);
            h1 = __reduce_add_sync(0xffffffffu, h1);
            eA = __reduce_add_sync(0xffffffffu, eA);
            eB = __reduce_add_sync(0xffffffffu, eB);
            const uint32_t word = lane < 4 ? h0 : lane < 8 ? h1 : lane < 12 ? eA : eB;
            const uint32_t v = lane < 16 ? (word >> (8 * (lane & 3))) & 0xFFu : 0u;
            if (v) {
                unsigned int* dst = lane < 8 ? &sm.cnt[lane]                                 // end-of-day status counts
                                  : lane < 12 ? &sm.ev[lane - 8]                             // EV_ADM .. EV_DAILY_QUAR
                                  : lane == 12 ? &sm.ev[EV_NEW_DIAG] : lane == 13 ? &sm.ev[EV_CHANGED]
                                  : lane == 14 ? &sm.cnt[8] : &sm.cnt[9];                    // susceptible / infectious at day start
                atomicAdd(dst, v);
            }
            const unsigned km = __match_any_sync(0xffffffffu, kbin);
            if (kbin >= 0 && lane == __ffs(km) - 1) atomicAdd(&sm.hist[kbin], (unsigned)__popc(km));
        }
        if (evm) {
            if (lane == 0) atomicAdd(&sm.ev[EV_EVENTS], (unsigned)__popc(evm));
            if (c.ev) {
                ev_base = __shfl_sync(0xffffffffu, ev_base, 0);
                DYS_ASSERT(c, ev_base >= 0 && ev_base + __popc(evm) <= c.n_pad, 13);
                if (infected_today) c.ev[ev_base + __popc(evm & lt_mask)] = make_int2((int)(c.lo + a), infector);
            }
        }
        if (cn) {
            // infectious tomorrow: queue the agent's row for the block's push of day + 1
            const bool f = (ibn & IB_INF) != 0;
            const unsigned fm = __ballot_sync(0xffffffffu, f);
            if (fm) {
                int base = 0;
                if (lane == 0) base = (int)atomicAdd(&sm.n_front, (unsigned)__popc(fm));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (f) {
                    DYS_ASSERT(c, base + __popc(fm & lt_mask) < FCAP, 12);
                sm.front[base + __popc(fm & lt_mask)] = rel | (ibn << 24);
                    prefetch_l2(cn->tptr + fmap.at(rel));
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// agent_body.  A block walks pieces of 128 agents in batches of <= 16 (two per warp), in two phases.
//   sweep  : 4 agents per lane: one 32-bit load of the status bytes, one of the attention bytes.  An agent to whom nothing
//            can happen today -- susceptible or long recovered, attention byte 0 -- only feeds packed counters.  Every other
//            agent is compacted (ballot + prefix popcount) into the block's work list, which accumulates over batches.
//   settle : once the list might not take another batch (or after the last batch) it is processed densely, one agent per
//            thread whatever piece it came from: B, C, applying D, the ordered rules of E, F, G, tomorrow's snapshot byte and
//            household flags; the next round's records are prefetched while a round is settled.  Agents that are infectious
//            tomorrow join the block's frontier, which is pushed (cn != null) right after the settle.
// Divergent per-agent logic therefore costs instructions in proportion to the agents that need it, not to the population,
// and the work of a batch is spread evenly over the block's warps however unevenly the epidemic is spread over the pieces.
// VARIANT 0 = k_agent (the last block to finish folds the statistics); 3 = k_days (one block folds after the grid barrier).
// `a_base` = first agent the block may meet (rel indices of the list and the frontier are relative to it).
// ---------------------------------------------------------------------------------------------------------------
template <int VARIANT, bool PEERS>
__device__ __forceinline__ void agent_body(const DevCtx& c, const int day, BlockScratch& sm, const DevCtx* cn, WorkSrc ws,
                                           const int64_t a_base, const bool record_cost = true)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid < HIST_LEN) sm.hist[tid] = 0;
    if (tid < EV_LEN) sm.ev[tid] = 0;
    if (tid < 10) sm.cnt[tid] = 0;
    if (tid == 0) {
        sm.n_list = 0; sm.n_front = 0;
        for (int p = 0; p < 8; ++p) sm.pub_cnt[p] = 0;
        for (int p = 0; p < 4; ++p) sm.push_acc[p] = 0;
        int n;
        const int64_t lo = take_result(take_issue(ws, 2 * WARPS), n);
        sm.b_lo[0] = lo; sm.b_n[0] = n;
        if (PEERS && c.n_wr > 1) count_pub(c, sm, lo, n);
    }
    if (cn && tid < 64) sm.pdf[tid] = c.pdf[tid];
    __syncthreads();
    const bool any_nd = c.df[DF_ANY_ND] != 0;
    // a recovered agent left the histogram (see k_validate_population for agents LOADED as recovered)
    const bool r_idle = c.recover >= HIST_LEN && c.hospital_recover >= HIST_LEN && day >= c.r_idle_day;
    const FrontMap fmap{c.lo - c.rt_lo + a_base};
    uint32_t n_idle_s = 0, n_idle_r = 0;
    constexpr int SLOTS = 2 * WARPS;              // pieces per batch

    // first batch's words
    int64_t b_lo = sm.b_lo[0];
    int b_n = sm.b_n[0];
    uint32_t sw[2], aw[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int sl = r * WARPS + wid;
        sw[r] = aw[r] = 0;
        if (sl < b_n) {
            const int64_t a0 = ((b_lo + sl) << 7) + lane * 4;
            DYS_ASSERT(c, b_lo >= 0 && a0 + 4 <= c.n_pad && a0 >= a_base, 10);
            sw[r] = *reinterpret_cast<const uint32_t*>(c.status + a0);
            aw[r] = *reinterpret_cast<const uint32_t*>(c.att + a0);
        }
    }
    for (int it = 0; b_n > 0; ++it) {
        // the batch after this one is taken now (its words are requested before this batch is settled); the answer of the
        // atomic is only looked at after the sweep
        unsigned long long taken = 0ull;
        if (tid == 0) taken = take_issue(ws, SLOTS);
        // ---- sweep ----
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int sl = r * WARPS + wid;
            const bool in = sl < b_n;
            const int64_t a0 = ((b_lo + sl) << 7) + lane * 4;
            const uint32_t rel0 = (uint32_t)(a0 - a_base);
            // four agents at once, byte-parallel: zb(x) has 0x80 in every byte of x that is zero (exact, no borrows)
            auto zb = [](uint32_t x) { return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu); };
            const uint32_t at0 = zb(aw[r]);                                           // attention byte == 0
            const uint32_t idle_s = zb(sw[r] ^ (0x01010101u * ST_S)) & at0;           // susceptible, nothing pending
            const uint32_t idle_r = r_idle ? zb(sw[r] ^ (0x01010101u * ST_R)) & at0 : 0u;
            n_idle_s += __popc(idle_s);
            n_idle_r += __popc(idle_r);
            const uint32_t wk = ~(idle_s | idle_r | zb(sw[r])) & 0x80808080u;         // 0x80 per agent that must be settled
            const uint32_t work = ((wk >> 7) | (wk >> 14) | (wk >> 21) | (wk >> 28)) & 0xFu;
            const int mine = __popc(work);
            if (c.piece_work && record_cost) {
                // what this piece costs a block (ns, measured with DYS_TRACE): ~50 for the sweep and ~65 per agent that is
                // settled (most of them are infectious and push a row tomorrow as well)
                const uint32_t cost = __reduce_add_sync(0xffffffffu, 65u * mine);
                if (in && lane == 0) c.piece_work[b_lo + sl] = 50u + cost;
            }
            if (in) {
                if (aw[r]) *reinterpret_cast<uint32_t*>(c.att + a0) = 0u;      // consumed: clean for the day after tomorrow
                // tomorrow's snapshot byte of an idle agent is 0; the settle phase overwrites the bytes of the agents it handles
                if (c.write_ib) *reinterpret_cast<uint32_t*>(c.ib_wr[0] + c.lo + a0) = 0u;
                if (PEERS && c.n_wr > 1) remote_ib_word(c, c.lo + a0, work);
            }
            const unsigned any = __ballot_sync(0xffffffffu, mine != 0);
            if (any) {
                int incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                int pos = 0;
                if (lane == 31) pos = (int)atomicAdd(&sm.n_list, (unsigned)incl);
                pos = __shfl_sync(0xffffffffu, pos, 31) + incl - mine;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if ((work >> j) & 1u) {
                        DYS_ASSERT(c, pos < LTOT && rel0 + j <= REL_MASK, 9);
                        sm.list[pos++] = (rel0 + j) | (((sw[r] >> (8 * j)) & 0xFu) << 24) | (((aw[r] >> (8 * j)) & 0x3u) << 28);
#if DYS_PREFETCH_REC
                        prefetch_l2(c.rec + a0 + j);      // the record is needed when the list is settled
#endif
                    }
            }
        }
        if (tid == 0) {
            int n;
            const int64_t lo = take_result(taken, n);
            sm.b_lo[(it + 1) & 1] = lo; sm.b_n[(it + 1) & 1] = n;
            if (PEERS && c.n_wr > 1) count_pub(c, sm, lo, n);
        }
        __syncthreads();
        // the next batch's words are requested before this batch is settled
        const int64_t nb_lo = sm.b_lo[(it + 1) & 1];
        const int nb_n = sm.b_n[(it + 1) & 1];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int sl = r * WARPS + wid;
            sw[r] = aw[r] = 0;
            if (sl < nb_n) {
                const int64_t a0 = ((nb_lo + sl) << 7) + lane * 4;
                DYS_ASSERT(c, nb_lo >= 0 && a0 + 4 <= c.n_pad && a0 >= a_base, 11);
                sw[r] = *reinterpret_cast<const uint32_t*>(c.status + a0);
                aw[r] = *reinterpret_cast<const uint32_t*>(c.att + a0);
            }
        }

        // ---- settle: once the list might not take another batch, or after the last batch.  A block in a quiet part of
        // the city sweeps all its pieces back to back and settles once ----
        const int n_list = (int)sm.n_list;
        const bool do_settle = n_list > LTOT - SLOTS * 128 || nb_n == 0;
        if (do_settle) {
            settle_list<VARIANT, PEERS>(c, day, sm, cn, a_base, any_nd);
            __syncthreads();      // the work list is free; the frontier count is final
            if (tid == 0) sm.n_list = 0;
            if (cn && sm.n_front) {      // (a settle can add as many rows as the frontier holds: push after each)
                push_frontier<false>(*cn, day + 1, sm, fmap.base);
                __syncthreads();
                if (tid == 0) sm.n_front = 0;
            }
        }
        __syncthreads();
        b_lo = nb_lo;
        b_n = nb_n;
    }
    {
        // idle agents: end-of-day status == day-start status
        const uint32_t tot_s = __reduce_add_sync(0xffffffffu, n_idle_s), tot_r = __reduce_add_sync(0xffffffffu, n_idle_r);
        if (lane == 0) {
            if (tot_s) { atomicAdd(&sm.cnt[1], tot_s); atomicAdd(&sm.cnt[8], tot_s); }
            if (tot_r) atomicAdd(&sm.cnt[6], tot_r);
        }
    }
    __syncthreads();
    if (PEERS && VARIANT == 3 && c.n_wr > 1) {
        // peer mode: every byte of the pieces this call took has been stored (here and in the peers).  Fence at system
        // scope, then tell each peer how many pieces it received.
        unsigned int any = 0;
        for (int p = 1; p < c.n_wr; ++p) any |= sm.pub_cnt[p];
        if (any) {
            __threadfence_system();
            __syncthreads();
            if (tid >= 1 && tid < c.n_wr && sm.pub_cnt[tid]) atomicAdd_system(c.arrive_at[tid], (unsigned long long)sm.pub_cnt[tid]);
        }
    }
    if (tid < 32) {
        // warp 0 publishes everything, fences, and takes the ticket
        const int slot = blockIdx.x & (N_SLOTS - 1);
        if (lane < 10 && sm.cnt[lane])
            atomicAdd(&c.part[(size_t)(RAW_END + lane) * N_SLOTS + slot], (unsigned long long)sm.cnt[lane]);
        if (lane < EV_LEN && sm.ev[lane]) {
            const int row = lane == EV_ADM ? ST_DAILY_HOSP : lane == EV_DEATH ? ST_DAILY_DEATHS : lane == EV_DAILY_INF ? ST_DAILY_INF
                          : lane == EV_DAILY_QUAR ? ST_DAILY_QUAR : lane == EV_NEW_DIAG ? ST_NEW_DIAG : lane == EV_EVENTS ? ST_N_EVENTS
                          : ST_CHANGED;
            atomicAdd(&c.part[(size_t)row * N_SLOTS + slot], (unsigned long long)sm.ev[lane]);
        }
        if (lane < HIST_LEN && sm.hist[lane])
            atomicAdd(&c.part[(size_t)(HIST0 + lane) * N_SLOTS + slot], (unsigned long long)sm.hist[lane]);
        if (VARIANT == 0) {
            __threadfence();
            __syncwarp();
            if (lane == 0) sm.last = (atomicAdd(c.ticket, 1u) == gridDim.x - 1);      // one ticket per block
        }
    }
    if (VARIANT != 0) return;
    __syncthreads();
    if (sm.last) {
        fold_stats(c.part, c.stats);      // the last block to finish
        if (tid == 0) *c.ticket = 0u;
    }
}

__global__ void __launch_bounds__(BLOCK) k_agent(const DevCtx c, const int day)
{
    __shared__ BlockScratch sm;
    __shared__ DevCtx s_c;               // (the phases are real functions: they take the context by address)
    if (threadIdx.x == 0) s_c = c;
    __syncthreads();
    const int64_t n_pieces = c.n_pad >> 7;
    const int64_t per = (n_pieces + gridDim.x - 1) / gridDim.x;
    const int64_t p0 = (int64_t)blockIdx.x * per < n_pieces ? (int64_t)blockIdx.x * per : n_pieces;
    const int64_t p1 = p0 + per < n_pieces ? p0 + per : n_pieces;
    agent_body<0, false>(s_c, day, sm, nullptr, WorkSrc{nullptr, 0, p0, p1}, p0 << 7);
}

// ---------------------------------------------------------------------------------------------------------------
// k_days: several consecutive days in ONE cooperative launch (persistent grid, all CTAs resident).
//   single GPU :  scan push of the first day, barrier; then per day  [agent(d) ; push(d+1) of the block's own
//                 frontier], ONE barrier.  (The last day of the window does not push: tomorrow's open flags arrive
//                 with the next call.)
//   peer mode  :  per day  wait for the peers' flags and push the HALO agents' rows (scan of their snapshot bytes), barrier,
//                 [agent(d) ; push(d+1) of the block's own frontier], barrier, publish the halo part of tomorrow's
//                 snapshot bytes to the peers and raise the flag.
// The per-day buffers come from DayPlan; each block keeps today's and tomorrow's view of the context in shared memory.
// ---------------------------------------------------------------------------------------------------------------
// peer mode: the owned slice of freshly written snapshot bytes (ib_wr[0]) is stored into the same place of every peer
// GPU's buffer (ib_wr[1..]) with coalesced 16-byte P2P stores over NVLink; the flag is raised afterwards (k_signal, or
// the tail of a k_days day) once every block's stores are fenced
// (pub_idx of n_pub blocks share the copy)
__device__ __forceinline__ void publish_body(const DevCtx& c, const int pub_idx, const int n_pub)
{
    // only the part of the owned slice that lies inside the peer's halo travels (pub_lo / pub_n, 16-byte granules)
    for (int p = 1; p < c.n_wr; ++p) {
        const uint4* src = reinterpret_cast<const uint4*>(c.ib_wr[0] + c.pub_lo[p]);
        uint4* dst = reinterpret_cast<uint4*>(c.ib_wr[p] + c.pub_lo[p]);
        const int64_t n16 = c.pub_n[p] >> 4;
        for (int64_t i = (int64_t)pub_idx * blockDim.x + threadIdx.x; i < n16; i += (int64_t)n_pub * blockDim.x) dst[i] = src[i];
    }
    __threadfence_system();
}

__global__ void __launch_bounds__(BLOCK) k_publish(const DevCtx c) { publish_body(c, (int)blockIdx.x, (int)gridDim.x); }

// Phase H on the device (contagious3.py:57-107, 368-385): the service block of k_days forms R of the day it just folded
// (compute_R, :29-42, in the reference's order of float operations) and decides the open flags THREE days ahead -- the
// decision for day t uses `Known R` of days t-1 and t-2, i.e. R of days t-1-diagnosis and t-2-diagnosis (:306-309), which are
// long known -- so a whole run needs no host round trip and no flag upload.
// SC_DIFF (contagious3.py:368-376, with the 'Vis_R' key of :371 read as the 'vis_R' that :319 writes): the same rule applied
// area by area, each area on the R of its own residents (the day's per-area histogram, complete after the barrier).
constexpr int SC_GRADUAL = 1, SC_ALL = 2, SC_EDU = 4, SC_REL = 8;      // scenario keywords; building class bits = 2, 4, 8
constexpr int SC_DIFF = 16;
struct LockdownPlan {
    int32_t mode;                                   // 0: the flags come from the host
    int32_t scen;                                   // SC_* bits
    const uint8_t* bld_class;                       // [B] 2: land use >= 3 (ALL), 4: EDU code, 8: REL code
    uint8_t* flags;                                 // [4][n_rep][B] ring by day & 3
    int32_t* n_closed;                              // [4][n_rep]
    double* r_hist;                                 // [16][n_rep] ring by day & 15: R of that day
    const int32_t* bld_area;                        // SC_DIFF: [B] the building's area (row of the per-area histogram)
    double* r_sa;                                   // SC_DIFF: [n_rep][16][S] ring by day & 15: per-area R of that day
    uint8_t* what_sa;                               // SC_DIFF: [n_rep][S] the decision per area (scratch of the service block)
    int32_t S;
};

constexpr int MAX_FUSED_DAYS = 14;
constexpr int PUBLISH_BLOCKS = 8;          // the halo part of a slice is small (one community per neighbour)
constexpr int HALO_BLOCKS = 16;            // blocks that push the halo agents' rows inside a k_days window
struct DayPlan {
    int32_t first_day, n_days;
    int32_t n_closed[MAX_FUSED_DAYS];
    const uint8_t* open[MAX_FUSED_DAYS];
    unsigned char* out[MAX_FUSED_DAYS];             // the day's output block
    size_t off_sa, off_bc, off_ev;                  // 0 = that part was not requested
    DayBuffers buf;
    int32_t want_sa, want_bc, want_ev;
    unsigned long long* trace;                      // developer tracing: [n_days][grid][8] globaltimer at phase boundaries, or null
    LockdownPlan ld;
    int64_t plan_max_pieces;                        // the equal-work cut is only planned up to this many pieces (beyond: equal split + stealing)
    unsigned long long epoch0;                      // peer mode: the bytes of first_day were the epoch0-th publication
    unsigned long long* work;                       // [grid] the blocks' work words (see WorkSrc), or null: no stealing
    int32_t test_hang_block;                        // tests (DYS_TEST_HANG_BLOCK): this block leaves before the first barrier; -1 = none
};

// the whole block copies a context (a single thread took ~950 instructions for it, at the start of every day, with
// everybody else waiting at the barrier behind it -- 6 % of the kernel's instructions in the ncu source view of r02p)
__device__ __forceinline__ void copy_ctx(DevCtx& dst, const DevCtx& src)
{
    static_assert(sizeof(DevCtx) % 4 == 0, "word copy");
    uint32_t* d = reinterpret_cast<uint32_t*>(&dst);
    const uint32_t* q = reinterpret_cast<const uint32_t*>(&src);
    for (int i = threadIdx.x; i < (int)(sizeof(DevCtx) / 4); i += blockDim.x) d[i] = q[i];
}

// one thread: turn a copy of the launch context into the view of day first_day + k
__device__ __forceinline__ void bind_plan_day(DevCtx& c, const DevCtx& c0, const DayPlan& plan, const int k, const int wb)
{
    bind_day(c, plan.buf, plan.first_day + k);
    c.open = plan.open[k];
    c.n_closed = plan.n_closed[k];
    c.open_stride = 0; c.n_closed_ptr = nullptr;
    if (plan.ld.mode) {      // (decided at least two grid barriers ago, see lockdown_service)
        const int R = c0.n_rep > 1 ? c0.n_rep : 1;
        const int slot = (plan.first_day + k) & 3;
        c.open = plan.ld.flags + (size_t)slot * R * c0.B;
        c.open_stride = c0.B;
        c.n_closed_ptr = plan.ld.n_closed + slot * R;
        c.n_closed = __ldcg(c.n_closed_ptr);
    }
    c.stats = reinterpret_cast<int64_t*>(plan.out[k]);
    c.sa_hist = plan.want_sa ? reinterpret_cast<int32_t*>(plan.out[k] + plan.off_sa) : nullptr;
    c.bcount = plan.want_bc ? reinterpret_cast<int32_t*>(plan.out[k] + plan.off_bc) : nullptr;
    c.ev = plan.want_ev ? reinterpret_cast<int2*>(plan.out[k] + plan.off_ev) : nullptr;
    // tomorrow's snapshot bytes are only scanned by a stand-alone push: every day in peer / two-phase mode, else only
    // after the last day of the window (the next call starts with a scan)
    c.write_ib = (c0.n_peers > 1 || c0.io_mat != nullptr || k == plan.n_days - 1) ? 1 : 0;
    c.epoch = plan.epoch0 + (unsigned long long)k;
    // this block's run of pieces on that day: its cut was finished at least a day ago (plan_cut runs three days ahead), so
    // fetching it here -- while the view is built, a day early for every day but the window's first -- costs no round trip later
    const int32_t* cut = plan.buf.bounds[(plan.first_day + k) % 3];
    if (cut && gridDim.x > 1 && wb >= 0) { c.cut_first = __ldcg(cut + wb); c.cut_last = __ldcg(cut + wb + 1); }
}

// The service block re-cuts the population for the day after tomorrow: contiguous runs of pieces with equal measured
// cost.  Piece q goes to block floor(prefix(q) * n_work / total); bounds[b] = the first piece of block b.  The piece
// costs are staged in the block's (otherwise idle) frontier + work-list shared memory, 8192 at a time, so that every
// global load is coalesced and all of them are in flight together; up to a million agents that is a single chunk and
// the total falls out of the same scan.
__device__ __forceinline__ void plan_cut(const uint32_t* __restrict__ work, int32_t* __restrict__ bounds, const int64_t n_pieces,
                                         const int n_work, BlockScratch& sm, unsigned long long* tr = nullptr)
{
    auto stamp = [&](int i) {
        if (tr && threadIdx.x == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            tr[i] = t;
        }
    };
    stamp(8);
    __shared__ unsigned long long s_sum[BLOCK];
    __shared__ unsigned long long s_carry, s_total;
    static_assert(offsetof(BlockScratch, list) == sizeof(uint32_t) * FCAP, "front and list are contiguous");
    const int tid = threadIdx.x;
    constexpr int CHUNK = FCAP + LTOT, PER = CHUNK / BLOCK;      // 32 consecutive pieces per thread and chunk
    uint32_t* const stage = sm.front;
    // thread t walks 32 consecutive words: XOR the word's low index with t so that the 32 lanes of a warp hit 32 banks
    auto at = [](int i) { return (i & ~31) | ((i & 31) ^ ((i >> 5) & 31)); };
    const bool single = n_pieces <= CHUNK;
    if (!single) {      // a first pass for the total
        unsigned long long s = 0;
        for (int64_t q = tid; q < n_pieces; q += BLOCK) s += __ldcg(work + q);
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((tid & 31) == 0) s_sum[tid >> 5] = s;
        __syncthreads();
        if (tid == 0) {
            unsigned long long t = 0;
            for (int w = 0; w < WARPS; ++w) t += s_sum[w];
            s_total = t;
        }
    }
    if (tid == 0) { s_carry = 0; bounds[0] = 0; bounds[n_work] = (int32_t)n_pieces; }
    __syncthreads();
    for (int64_t c0 = 0; c0 < n_pieces; c0 += CHUNK) {
        {
            uint32_t v[PER];                                 // all of the thread's loads are issued before the first store
#pragma unroll
            for (int j = 0; j < PER; ++j) v[j] = c0 + tid + j * BLOCK < n_pieces ? __ldcg(work + c0 + tid + j * BLOCK) : 0u;
#pragma unroll
            for (int j = 0; j < PER; ++j) stage[at(tid + j * BLOCK)] = v[j];
        }
        __syncthreads();
        stamp(9);
        unsigned long long mine = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) mine += stage[at(tid * PER + k)];
        s_sum[tid] = mine;
        __syncthreads();
        if (tid < 32) {      // exclusive scan of the 256 thread sums: 8 per lane, then a warp scan
            unsigned long long v[8], run = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) { v[k] = s_sum[tid * 8 + k]; run += v[k]; }
            unsigned long long incl = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
                if (tid >= o) incl += t;
            }
            unsigned long long ex = s_carry + incl - run;
#pragma unroll
            for (int k = 0; k < 8; ++k) { s_sum[tid * 8 + k] = ex; ex += v[k]; }
            __syncwarp();
            if (tid == 31) { s_carry += incl; if (single) s_total = incl; }
        }
        __syncthreads();
        stamp(10);
        // floor(before * n_work / total) with one division per thread instead of one per piece: (before * M) >> 32 with
        // M = floor(2^32 * n_work / total).  Any monotone function of `before` that every thread evaluates identically
        // gives a valid cut.  (total < 2^40 and n_work < 2^11: no overflow.)
        const unsigned long long M = (((unsigned long long)n_work) << 32) / (s_total ? s_total : 1ull);
        auto block_of = [&](unsigned long long before) {
            const int b = (int)((before * M) >> 32);
            return b < n_work - 1 ? b : n_work - 1;
        };
        unsigned long long cum = s_sum[tid];                 // cost of all pieces before this thread's first
        const int64_t q0 = c0 + tid * PER;
        int prev = 0;
        if (q0 > 0 && q0 < n_pieces) prev = block_of(cum - (tid ? stage[at(tid * PER - 1)] : __ldcg(work + q0 - 1)));
#pragma unroll 4
        for (int k = 0; k < PER; ++k) {
            const int64_t q = q0 + k;
            if (q < n_pieces) {
                const int b = block_of(cum);
                for (int j = prev + 1; j <= b; ++j) bounds[j] = (int32_t)q;
                prev = b;
                cum += stage[at(tid * PER + k)];
                if (q == n_pieces - 1)
                    for (int j = prev + 1; j < n_work; ++j) bounds[j] = (int32_t)n_pieces;      // blocks left without pieces
            }
        }
        __syncthreads();
    }
}

// after the fold of `day` for replica r: R(day) into the history, 'Closed buildings' / R / Known R into the day's stats block
// (slots 28-30), and the flags of day + 3
__device__ __noinline__ void lockdown_service(const LockdownPlan ld, const int64_t B, const int recover, const int diagnosis,
                                              const double* __restrict__ pdf, const uint64_t seed, const int day, const int r,
                                              const int n_rep, int64_t* __restrict__ stats, const int32_t* __restrict__ sa_hist)
{
    __shared__ double s_vr[2];
    __shared__ int s_closed;
    const int tid = threadIdx.x;
    const int t = day + 3;
    if (tid == 0) {
        double sum_I = 0.0;                                                                     // compute_R, :29-42
        for (int i = 1; i < recover; ++i) sum_I = __dadd_rn(sum_I, __dmul_rn((double)stats[HIST0 + i], pdf[i]));
        const double R = sum_I > 0.0 ? __ddiv_rn((double)stats[HIST0], sum_I) : 0.0;
        ld.r_hist[(day & 15) * n_rep + r] = R;
        auto known = [&](const int d) { return d > diagnosis ? ld.r_hist[((d - diagnosis) & 15) * n_rep + r] : 0.0; };   // :306-309
        s_vr[0] = known(t - 1);                                                                 // vR of the decision made on day t-1
        s_vr[1] = (t - 1 != 0) ? known(t - 2) : 0.0;                                            // prevVR, :379-384
        stats[28] = ld.n_closed[(day & 3) * n_rep + r];
        stats[29] = __double_as_longlong(R);
        stats[30] = __double_as_longlong(known(day));
        s_closed = 0;
    }
    __syncthreads();
    enum { OPEN, FULL, HALF, KEEP };
    auto decide = [&](const double vR, const double prevVR) {                                   // building_lockdown, :57-107
        if (ld.scen & SC_GRADUAL) {
            if (vR > 1.0 && vR < 2.0) return (prevVR <= 1.0 || prevVR >= 2.0) ? HALF : KEEP;
            return vR >= 2.0 ? FULL : OPEN;
        }
        return vR > 1.0 ? FULL : OPEN;
    };
    const bool diff = (ld.scen & SC_DIFF) != 0;
    const uint8_t* what_sa = nullptr;
    int what;
    if (diff) {
        // per area: R of the area's residents today (sas_R, :282-291) into the area's history, then the decision from the
        // area's visible R of days t-1 and t-2 (sas_vis_R, :311-318; :369-376).  t-1-diagnosis <= day: already in the ring.
        const int S = ld.S;
        double* ring = ld.r_sa + (size_t)r * 16 * S;
        uint8_t* w_out = ld.what_sa + (size_t)r * S;
        int any = 0;
        for (int s = tid; s < S; s += blockDim.x) {
            const int32_t* hs = sa_hist + (size_t)s * HIST_LEN;
            double sum_I = 0.0;
            for (int i = 1; i < recover; ++i) sum_I = __dadd_rn(sum_I, __dmul_rn((double)__ldcg(hs + i), pdf[i]));
            ring[(size_t)(day & 15) * S + s] = sum_I > 0.0 ? __ddiv_rn((double)__ldcg(hs), sum_I) : 0.0;
            auto known_s = [&](const int d) { return d > diagnosis ? ring[(size_t)((d - diagnosis) & 15) * S + s] : 0.0; };
            const int w = decide(known_s(t - 1), (t - 1 != 0) ? known_s(t - 2) : 0.0);
            w_out[s] = (uint8_t)w;
            any |= w != OPEN;
        }
        what = __syncthreads_or(any) ? KEEP : OPEN;          // (anything but OPEN: some area decides per building below)
        what_sa = w_out;
    } else
        what = decide(s_vr[0], s_vr[1]);
    const uint32_t classes = (ld.scen & SC_ALL) ? (uint32_t)SC_ALL : (uint32_t)(ld.scen & (SC_EDU | SC_REL));
    uint8_t* dst = ld.flags + ((size_t)(t & 3) * n_rep + r) * B;
    const uint8_t* prev = ld.flags + ((size_t)((t - 1) & 3) * n_rep + r) * B;
    int* const slot_closed = ld.n_closed + (t & 3) * n_rep + r;
    if (what == OPEN && *slot_closed == 0) { __syncthreads(); return; }       // (the slot still holds an all-open vector)
    int closed = 0;
    for (int64_t b = tid; b < B; b += blockDim.x) {
        uint8_t f = 1;
        const int w = diff ? (int)what_sa[__ldg(ld.bld_area + b)] : what;
        const bool in_class = (__ldg(ld.bld_class + b) & classes) != 0;
        if (w == FULL) f = in_class ? 0 : 1;
        // (the reference draws exactly half of each class with np.random.choice; the device policy closes each building of
        // the class with probability 1/2 on a keyed draw -- host twin: host.building_lockdown(keyed=...))
        else if (w == HALF) f = (in_class && keyed_uniform(seed, (uint32_t)(t - 1), STREAM_LOCKDOWN, (uint32_t)b, 0u) < 0.5) ? 0 : 1;
        else if (w == KEEP) f = prev[b];
        dst[b] = f;
        closed += f == 0;
    }
    closed = __reduce_add_sync(0xffffffffu, closed);
    if ((tid & 31) == 0 && closed) atomicAdd(&s_closed, closed);
    __syncthreads();
    if (tid == 0) *slot_closed = s_closed;
    __syncthreads();
}

// All CTAs are resident (cooperative launch): one arrival per CTA on a monotone counter, no reset needed.
__device__ __forceinline__ bool grid_barrier(const DevCtx& c, unsigned int* counter, unsigned int& epoch)
{
    __shared__ int s_abort;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        ++epoch;
        const unsigned int target = epoch * gridDim.x;
        s_abort = spin_until(c, [&]() { return *reinterpret_cast<volatile unsigned int*>(counter) >= target; }) ? 1 : 0;
        __threadfence();
    }
    __syncthreads();
    return s_abort != 0;
}

// PEERS = false: the single-GPU kernel carries none of the exchange code (16 % faster than the generic one at 1 M agents)
// MULTI = false: one replica and runs short enough for the planned cut alone -- no replica views, no work words, no stealing
template <bool PEERS, bool MULTI>
__global__ void __launch_bounds__(BLOCK, 4) k_days(const __grid_constant__ DevCtx c0, const __grid_constant__ DayPlan plan,
                                                   unsigned int* barrier_counter)
{
    __shared__ BlockScratch sm;
    __shared__ DevCtx s_c[2];            // today's and tomorrow's view of the context (read like kernel parameters)
    __shared__ DevCtx s_r[2];            // replicas: the same two views shifted to the replica the block is working on
    unsigned int epoch = 0;
    const bool peers = PEERS && c0.n_peers > 1;
    const int n_rep = MULTI && c0.n_rep > 1 ? c0.n_rep : 1;
    // io_mat is one buffer for the latest day: the push of day + 1 must not clear it under the agent pass of day
    const bool two_phase = c0.io_mat != nullptr;
    // Block 0 owns no agents: it folds each day's statistics and re-cuts the population after the barrier while the
    // others are already in the next day.  (The FIRST block, because the warp schedulers favour the oldest warps of an
    // SM: as the youngest block it was starved by its three busy neighbours and became the one everybody waited for.)
    // Peer mode with a boundary split: HALO_BLOCKS more blocks own no agents either -- they wait for the partners' bytes and
    // push the halo agents' rows for tomorrow while the work blocks are busy with the bulk of today (the halo is ~1 % of a
    // day's push work: one block alone would be the last to reach the barrier).
    // (DYS_BOUNDARY_FIRST, off: measured at N = 2, 1 M agents per GPU -- 44.9 us per day with the split and 16 halo blocks,
    // 46.8 with one, against 37.5 when the first 16 work blocks push the halo rows after their own work: the extra
    // agent_body call + system fence of the boundary phase costs more than the exchange latency it hides)
    const int n_service = (DYS_BOUNDARY_FIRST && PEERS && peers && c0.bnd_ok && gridDim.x > 4 * HALO_BLOCKS) ? 1 + HALO_BLOCKS : 1;
    const bool service = gridDim.x > 1 && blockIdx.x == 0;
    const bool halo_block = n_service > 1 && blockIdx.x >= 1 && (int)blockIdx.x < n_service;
    const int n_work = gridDim.x > 1 ? (int)gridDim.x - n_service : 1;
    const int wb = gridDim.x > 1 ? (int)blockIdx.x - n_service : 0;
    const int64_t n_pieces = c0.n_pad >> 7, n_pieces_all = n_pieces * n_rep;
    // long runs are balanced dynamically (work words + stealing), short ones by the measured cut alone
    const bool stealing = MULTI && plan.work != nullptr && gridDim.x > 1 && n_pieces_all > 32ll * n_work;
    copy_ctx(s_c[0], c0);
    __syncthreads();
    if (threadIdx.x == 0) bind_plan_day(s_c[0], c0, plan, 0, wb);
    __syncthreads();
    if ((int)blockIdx.x == plan.test_hang_block) return;      // (tests: the others must time out at the barrier, not hang)
    // the view of replica r of a day: with one replica the day's view itself, no copy
    auto replica_view = [&](const DevCtx& day_view, const int slot, const int r) -> const DevCtx& {
        if (!MULTI || n_rep == 1) return day_view;
        __syncthreads();                       // (everybody is done with the previous occupant of the slot)
        copy_ctx(s_r[slot], day_view);
        __syncthreads();
        if (threadIdx.x == 0) shift_replica(s_r[slot], r);
        __syncthreads();
        return s_r[slot];
    };
    for (int k = 0; k < plan.n_days; ++k) {
        const int day = plan.first_day + k;
        const DevCtx& c = s_c[k & 1];
        const bool local_push = c0.io_mat == nullptr && k + 1 < plan.n_days;
        if (k + 1 < plan.n_days) {
            copy_ctx(s_c[(k + 1) & 1], c0);
            __syncthreads();
            if (threadIdx.x == 0) bind_plan_day(s_c[(k + 1) & 1], c0, plan, k + 1, wb);
        }
        __syncthreads();
        unsigned long long* tr = plan.trace ? plan.trace + ((size_t)k * gridDim.x + blockIdx.x) * TRACE_SLOTS : nullptr;
        auto stamp = [&](int i) {
            if (tr && threadIdx.x == 0) tr[i] = global_ns();
        };
        stamp(0);
        if (k == 0 || two_phase) {
            for (int r = 0; r < n_rep; ++r) {
                const DevCtx& cr = replica_view(c, 0, r);
                if (push_scan_body<false, PEERS>(cr, day, sm, n_work, wb)) return;
                __syncthreads();
            }
            stamp(1);
            if (grid_barrier(c, barrier_counter, epoch)) return;
        }
        stamp(2);
        const DevCtx* cn = local_push ? &s_c[(k + 1) & 1] : nullptr;
        if (cn)
            for (int r = 0; r < n_rep; ++r) push_clear_duties(*cn, r);
        {
            // The block's run of pieces in the unified piece space [replica][piece] -- today's cut of equal measured work,
            // or an equal share -- one agent_body per replica the run touches.  Long runs are consumed through the
            // block's work word, from which blocks that are done take batches as well (work stealing): the epidemic is
            // never evenly spread, and no cost model predicts an SM's speed.
            int64_t P0, P1;
            if (wb < 0 || wb >= n_work) P0 = P1 = 0;
            else if (c.cut_first >= 0) { P0 = c.cut_first; P1 = c.cut_last; }
            else {
                const int64_t per = (n_pieces_all + n_work - 1) / n_work;
                P0 = wb * per < n_pieces_all ? wb * per : n_pieces_all;
                P1 = P0 + per < n_pieces_all ? P0 + per : n_pieces_all;
            }
            int bnd_seg = 0;                    // boundary pieces this block settles first: up to two ranges
            int64_t bnd_lo0 = 0, bnd_hi0 = 0, bnd_lo1 = 0, bnd_hi1 = 0;
            if (PEERS && n_service > 1 && wb >= 0 && wb < n_work) {
                // the boundary pieces (prefix [0, a) and suffix [b, n) of the owned range) are dealt to ALL work blocks and
                // settled before anything else; the own run covers the interior only (their recorded cost stays 0, so the
                // planned cut ignores them)
                const int64_t a = c.bnd_a, b = c.bnd_b, nb = a + (n_pieces - b);
                const int64_t q = (nb + n_work - 1) / n_work;
                const int64_t i0 = (int64_t)wb * q < nb ? (int64_t)wb * q : nb, i1 = i0 + q < nb ? i0 + q : nb;
                if (i0 < a) { bnd_lo0 = i0; bnd_hi0 = i1 < a ? i1 : a; bnd_seg |= 1; }
                if (i1 > a) { bnd_lo1 = b + ((i0 > a ? i0 : a) - a); bnd_hi1 = b + (i1 - a); bnd_seg |= 2; }
                P0 = P0 > a ? P0 : a;
                P1 = P1 < b ? P1 : b;
            }
            // (one loop, one call site of agent_body: boundary pieces, then the segments of the own run, then stolen batches)
            int64_t P = P0;
            bool announced = false;
            const unsigned int all_done = (unsigned)(k + 1) * (unsigned)n_work;
            for (int round = 0;;) {
                int r;
                WorkSrc ws;
                int64_t a_base;
                bool record = true;
                if (PEERS && bnd_seg != 0) {
                    const bool first = (bnd_seg & 1) != 0;
                    const int64_t lo = first ? bnd_lo0 : bnd_lo1, hi = first ? bnd_hi0 : bnd_hi1;
                    bnd_seg &= first ? ~1 : ~2;
                    r = 0;
                    a_base = lo << 7;
                    ws = WorkSrc{nullptr, 0, lo, hi};
                    record = false;
                } else if (P < P1) {
                    r = (int)(P / n_pieces);
                    const int64_t p0 = P - (int64_t)r * n_pieces;
                    const int64_t p1 = P1 - (int64_t)r * n_pieces < n_pieces ? P1 - (int64_t)r * n_pieces : n_pieces;
                    a_base = p0 << 7;
                    if (stealing && p1 - p0 > 32) {
                        if (threadIdx.x == 0) atomicExch(&plan.work[wb], pack_work(P, P + (p1 - p0)));
                        ws = WorkSrc{&plan.work[wb], (int64_t)r * n_pieces, 0, 0};
                    } else {
                        ws = WorkSrc{nullptr, 0, p0, p1};
                    }
                    P += p1 - p0;
                } else {
                    if (!stealing || wb < 0 || wb >= n_work) break;
                    // done with the own run: help whoever is not.  Warp 0 reads 32 blocks' work words at a time and takes a
                    // batch from the one with the most pieces left; when every block has exhausted its own run there is
                    // nothing left to take.
                    if (!announced) {
                        announced = true;
                        if (threadIdx.x == 0) { __threadfence(); atomicAdd(barrier_counter + 2, 1u); }
                    }
                    if (threadIdx.x < 32) {
                        const int lane = threadIdx.x;
                        const int v = (int)(((long long)wb + 1 + lane + 32ll * round) % n_work);
                        const unsigned long long w = __ldcg(plan.work + v);
                        const uint32_t nx = (uint32_t)w, en = (uint32_t)(w >> 32);
                        const uint32_t rem = (v != wb && en > nx) ? en - nx : 0u;
                        const uint32_t best = __reduce_max_sync(0xffffffffu, rem);
                        const unsigned who = __ballot_sync(0xffffffffu, rem == best && best > 0u);
                        if (best > 0u && lane == __ffs(who) - 1) {
                            WorkSrc vs{plan.work + v, 0, 0, 0};
                            int n = 0;
                            sm.b_lo[0] = take_result(take_issue(vs, WARPS), n);      // unified piece index (rep_off = 0)
                            sm.b_n[0] = n;
                        }
                        if (lane == 0) {
                            if (best == 0u) sm.b_n[0] = 0;
                            sm.b_n[1] = (int)(__ldcg(barrier_counter + 2) >= all_done);
                        }
                    }
                    ++round;
                    __syncthreads();
                    const int n = sm.b_n[0];
                    const int64_t Pst = sm.b_lo[0];
                    const bool finished = sm.b_n[1] != 0;
                    __syncthreads();
                    if (n == 0) {
                        if (finished) break;
                        __nanosleep(400);          // (nothing to take right now: do not hammer the work words and the issue slots)
                        continue;
                    }
                    r = (int)(Pst / n_pieces);
                    const int64_t p0 = Pst - (int64_t)r * n_pieces;
                    a_base = p0 << 7;
                    ws = WorkSrc{nullptr, 0, p0, p0 + n};
                }
                const DevCtx& cr = replica_view(c, 0, r);
                const DevCtx* cnr = cn ? &replica_view(*cn, 1, r) : nullptr;
                agent_body<3, PEERS>(cr, day, sm, cnr, ws, a_base, record);
                if (cnr) publish_push_counters(*cnr, sm);
                __syncthreads();
            }
        }
        if (PEERS && peers) {
            // One token per day and partner (the partners stay within a day of each other even if no bytes travel one way),
            // then the push of the HALO agents' rows for tomorrow: their bytes arrive from the peers as those finish today's
            // agent pass.  A few blocks share it -- the halo is a community or two -- and everybody meets at the barrier.
            if ((n_service > 1 ? blockIdx.x == 1 : wb == 0) && threadIdx.x >= 1 && (int)threadIdx.x < c.n_wr)
                atomicAdd_system(c.arrive_at[threadIdx.x], 1ull);
            const int n_halo = HALO_BLOCKS < n_work ? HALO_BLOCKS : n_work;
            const int hb = n_service > 1 ? (halo_block ? (int)blockIdx.x - 1 : -1) : wb;
            if (cn && hb >= 0 && hb < n_halo) {
                if (push_scan_body<true, true>(*cn, day + 1, sm, n_halo, hb)) return;
            }
        }
        stamp(3);
        if (grid_barrier(c, barrier_counter, epoch)) return;
        stamp(4);
        if (service || gridDim.x == 1) {                      // every block's partial sums are in; the others move on
            for (int r = 0; r < n_rep; ++r) {
                int64_t* st = reinterpret_cast<int64_t*>(reinterpret_cast<unsigned char*>(c.stats) + (size_t)r * c.out_stride);
                fold_stats(c.part + (size_t)r * N_STATS * N_SLOTS, st);
                if (plan.ld.mode)
                    lockdown_service(plan.ld, c.B, c.recover, c.diagnosis, c.pdf, c.rep_params ? c.rep_params[r].seed : c.seed, day, r, n_rep, st,
                                     c.sa_hist ? reinterpret_cast<const int32_t*>(reinterpret_cast<const unsigned char*>(c.sa_hist) + (size_t)r * c.out_stride) : nullptr);
            }
            // today's piece costs are complete: cut the population for the day after tomorrow (same parity as today)
            // (two days out of four: the load shifts over weeks, and a block that plans is late for the next barrier on a
            // quiet day.  The cut made after day d is the one of day d + 3.)
            if (c.piece_work && gridDim.x > 1 && (day & 2) == 0 && n_pieces_all <= plan.plan_max_pieces)
                plan_cut(c.piece_work, c.bounds_plan, n_pieces_all, n_work, sm, tr);
            stamp(11);
        }
        __syncthreads();      // the fold (and tomorrow's bind) are done with today's view
    }
}

// phase A alone (contagious3.py:181-202): the day-start snapshot byte of every owned agent and the households that see
// a diagnosis on `day` (raised into the day's own attention bytes / quirk flags / flag block, which the host cleared).
__global__ void __launch_bounds__(BLOCK) k_snapshot(const DevCtx c, const int day)
{
    const int64_t a0 = ((int64_t)blockIdx.x * BLOCK + threadIdx.x) * 4;
    if (a0 >= c.n_pad) return;
    const uchar4 sv = *reinterpret_cast<const uchar4*>(c.status + a0);
    const uint32_t st[4] = {sv.x, sv.y, sv.z, sv.w};
    uint8_t ibn[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const AgentRec r = c.rec[a0 + j];
        ibn[j] = infectious_byte(st[j], day, r.t_inf);
        if (diagnosed_on(c, c.u_adm, a0 + j, st[j], r.t_inf, day, r.p_adm)) raise_household(c, a0 + j, r.hh_slot, r.hh_size, c.att, c.qflag, c.df);
    }
    // (k_snapshot of `day` is bound like day - 1's agent pass would be: ib_wr[0] is the buffer of `day`)
    *reinterpret_cast<uchar4*>(c.ib_wr[0] + c.lo + a0) = make_uchar4(ibn[0], ibn[1], ibn[2], ibn[3]);
}

// peer mode, stand-alone launches (k_snapshot / k_agent + k_publish): every partner receives the day's units at once
__global__ void k_signal(const DevCtx c)
{
    if (threadIdx.x >= 1 && (int)threadIdx.x < c.n_wr) {
        __threadfence_system();
        atomicAdd_system(c.arrive_at[threadIdx.x], (unsigned long long)c.units_out[threadIdx.x]);
    }
}

}  // namespace dys
