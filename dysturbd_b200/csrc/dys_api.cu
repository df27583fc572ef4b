// C ABI of the B200-native epidemic step (include/dysturbd.h).  Host-side bookkeeping only: owns the device
// device arrays, launches the kernels of dys_day.cuh / dys_load.cuh on one stream, and moves the small per-day
// integer outputs to pinned host memory.  No CPU implementation of any phase lives here.
#include "../../include/dysturbd.h"

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <cmath>
#include <ctime>
#include <string>
#include <vector>

#include "dys_day.cuh"
#include "dys_load.cuh"
#include "dys_routine.cuh"
#include "dys_network.cuh"

using namespace dys;

static_assert(DYS_N_STATS == N_STATS && DYS_ST_HIST0 == HIST0 && DYS_HIST_LEN == HIST_LEN, "header/kernel mismatch");
static_assert(DYS_ST_CHANGED == ST_CHANGED && DYS_ST_N_EVENTS == ST_N_EVENTS, "header/kernel mismatch");

static thread_local std::string g_create_error;

struct dys_handle {
    dys_params p{};
    DevCtx c{};
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // the day's output block leaves on its own stream, under the next day's kernels
    cudaEvent_t day_done[DYS_OUT_RING] = {};        // per ring slot: the day's kernels finished writing the block
    // per-day open flags of a fused window: two staging rings by window parity, uploaded on their own stream so that the
    // flags of window w + 1 travel while window w runs
    uint8_t* d_open_ring[2] = {};                   // [MAX_FUSED_DAYS][B]
    uint8_t* h_open_ring[2] = {};                   // pinned staging of the same
    cudaStream_t up_stream = nullptr;
    cudaEvent_t up_done[2] = {}, ring_free[2] = {};
    bool ring_used[2] = {false, false};
    int64_t windows = 0;
    int R = 1;                                      // replicas
    int sel = -1;                                   // dys_select_replica
    std::vector<RepParam> rep;                      // host copy of the per-replica parameters
    RepParam* d_rep = nullptr;
    LockdownPlan ld{};                              // device lockdown policy (dys_set_lockdown)
    volatile int32_t* h_abort = nullptr;            // pinned + mapped: the kernels poll it in their spin loops
    int64_t h2d_bytes = 0;                          // flag bytes uploaded so far (bench accounting)
    unsigned int* d_barrier = nullptr;              // grid barrier counter of k_days, publish ticket, blocks done with their run
    unsigned long long* d_work = nullptr;           // [grid] work words of k_days (dynamic batches, stealing)
    int days_grid = 0;                              // cooperative grid of k_days (0 = not available)
    std::vector<void*> allocs;
    int64_t dev_bytes = 0;
    std::string err;
    bool have_pop = false, have_graph = false, begun = false;
    int64_t nnz = 0, nnz_kept = 0;
    int32_t last_day = -1;
    int64_t steps = 0;
    // one device output block [stats | sa_hist | bcount | events] per ring slot and the pinned host ring it is copied into
    unsigned char* d_out[DYS_OUT_RING] = {};          // device output blocks, one per ring slot (slot = step % ring)
    size_t off_sa = 0, off_bc = 0, off_ev = 0, out_copy_bytes = 0, out_dev_bytes = 0;
    int64_t ev_inline = 0;
    unsigned char* h_out = nullptr;                       // [DYS_OUT_RING][out_copy_bytes] pinned
    int32_t out_day[DYS_OUT_RING];
    int64_t out_step[DYS_OUT_RING];
    cudaEvent_t out_ev[DYS_OUT_RING];
    int out_ev_of[DYS_OUT_RING];                          // the slot whose event covers this slot's copy (a window ships as one)
    std::vector<int64_t> hist_stats;                      // [DYS_STATS_RING][R][N_STATS] stats of days that left the output ring
    int32_t hist_day[DYS_STATS_RING];
    uint8_t* h_open = nullptr;                            // pinned shadow of the open flags
    int32_t* h_small = nullptr;      // pinned scratch: [0] validation flags, [1] r_idle_day
    int32_t* d_bad = nullptr;
    DayBuffers buf{};                // attention bytes, mailboxes, quirk flags, flag blocks, partial sums: rings indexed by day
    int32_t ib_day = -1;             // the day whose day-start snapshot bytes are complete in buffer ib_day & 1
    uint8_t* d_ib = nullptr;         // [2][ib_stride] snapshot bytes of all agents, double buffered by day parity
    int64_t ib_stride = 0;
    unsigned long long* d_arrive = nullptr;   // [8] peer mode: units (pieces + tokens) received from rank r on this GPU
    unsigned long long pub_epoch = 0;         // publications of snapshot bytes so far (all ranks step the same days)
    uint8_t* peer_ib[8] = {};        // peer mode: every rank's d_ib (own included)
    unsigned long long* peer_arrive[8] = {};
    int n_peers = 1;
    int64_t hh_lo = 0, bld_lo = 0, area_lo = 0;   // first owned household / building / area
    int64_t peer_halo[8][2] = {};    // every rank's halo [lo, hi) in global agents
    int64_t peer_own[8][2] = {};     // every rank's owned range
    uint32_t partner_mask = 0;       // ranks this rank exchanges bytes with (either direction)
    int push_grid = 0;               // persistent grid of k_push
    int last_slot = 0;
    bool upload_always = false;       // DYS_UPLOAD_ALWAYS=1: every offered flag vector travels (bench accounting of host inputs)
    bool no_steal = false;
    bool test_hooks = false;          // DYS_TEST_HANG_BLOCK was set when the handle was created
    int64_t plan_max = 16384;
    bool no_fuse = false;            // DYS_NO_FUSE=1 in the environment: dys_step_many launches two kernels a day
    const char* trace_path = nullptr;          // DYS_TRACE=<file> (developer): per-block phase timestamps of every fused window
    unsigned long long* d_trace = nullptr;
    double* d_pair_prob = nullptr;
    double *d_u_adm = nullptr, *d_u_death = nullptr, *d_u_pair = nullptr;   // staged injected draws
    // kernel timing
    std::vector<cudaEvent_t> tev;    // per step: 4 events (before K1, after K1 / before K2, after K2, after K3)
    int64_t t_begin = 0, t_end = 0;  // steps [t_begin, t_end) hold unread timings
    static constexpr int64_t T_RING = 4096;
};

static int fail(dys_handle* h, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}

#define CK(h, call)                                                                                         \
    do {                                                                                                    \
        cudaError_t e_ = (call);                                                                            \
        if (e_ != cudaSuccess) return fail(h, DYS_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

template <typename T>
static int dev_alloc(dys_handle* h, T** out, int64_t count, bool zero = true)
{
    void* p = nullptr;
    const size_t bytes = (size_t)(count > 0 ? count : 1) * sizeof(T);
    CK(h, cudaMalloc(&p, bytes));
    if (zero) CK(h, cudaMemsetAsync(p, 0, bytes, h->stream));
    h->allocs.push_back(p);
    h->dev_bytes += (int64_t)bytes;
    *out = (T*)p;
    return DYS_OK;
}

// release one tracked allocation early (a re-loaded network / quirk table)
static void dev_free(dys_handle* h, const void* p)
{
    if (!p) return;
    for (size_t i = 0; i < h->allocs.size(); ++i)
        if (h->allocs[i] == p) { h->allocs.erase(h->allocs.begin() + (long)i); break; }
    cudaFree(const_cast<void*>(p));
}

static int check_abort(dys_handle* h)
{
    if (h->h_abort && *h->h_abort >= 1000)
        return fail(h, DYS_ERR_ABORTED, "internal consistency check %d failed (checked build, see DYS_ASSERT in csrc/)", *h->h_abort - 1000);
    if (h->h_abort && *h->h_abort)
        return fail(h, DYS_ERR_ABORTED, "a launch was abandoned (%s); reload the state with dys_set_state",
                    *h->h_abort == 2 ? "a grid barrier or peer wait timed out: DYS_SPIN_TIMEOUT_MS" : "dys_abort");
    return DYS_OK;
}

static int copy_in(dys_handle* h, void* dst, const void* src, size_t bytes)
{
    if (bytes == 0) return DYS_OK;
    CK(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, h->stream));
    return DYS_OK;
}

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline unsigned grid_for(int64_t n, int per_block) { return (unsigned)((n + per_block - 1) / per_block); }

// ---- network construction: handle-free calls, host pointers in and out ----
namespace {
struct DevTmp {                      // device scratch of one call, freed on every exit path
    std::vector<void*> ptrs;
    cudaError_t e = cudaSuccess;
    template <typename T>
    T* get(size_t count)
    {
        void* d = nullptr;
        if (e == cudaSuccess) e = cudaMalloc(&d, (count ? count : 1) * sizeof(T));
        if (e == cudaSuccess) ptrs.push_back(d);
        return (T*)d;
    }
    template <typename T>
    T* up(const T* src, size_t count)
    {
        T* d = get<T>(count);
        if (e == cudaSuccess && count) e = cudaMemcpy(d, src, count * sizeof(T), cudaMemcpyHostToDevice);
        return d;
    }
    ~DevTmp() { for (void* p : ptrs) cudaFree(p); }
};

int pick_device(int device, const char* who)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(nullptr, DYS_ERR_CUDA, "%s: no CUDA device: this library has no CPU fallback", who);
    if (device < 0 || device >= ndev) return fail(nullptr, DYS_ERR_INVALID_ARG, "%s: device %d out of range", who, device);
    if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, DYS_ERR_CUDA, "%s: cudaSetDevice failed", who);
    return DYS_OK;
}
}  // namespace

extern "C" int dys_similarity_edges(int device, int64_t n, const double* age, const double* income, const double* x, const double* y,
                                    const int64_t* household, double w_a, double w_i, double w_d, double min_prob,
                                    int64_t* indptr /* [n+1] out */, int64_t capacity, int32_t* col, double* prob, double* maxes /* [3] out, nullable */)
{
    const char* who = "dys_similarity_edges";
    if (n <= 0 || !age || !income || !x || !y || !household || !indptr) return fail(nullptr, DYS_ERR_INVALID_ARG, "%s: null or empty argument", who);
    if (n > 0x7fffffffll) return fail(nullptr, DYS_ERR_RANGE, "%s: more than 2^31 - 1 agents", who);
    for (int64_t i = 0; i < n; ++i)
        if (!std::isfinite(age[i]) || !std::isfinite(income[i]) || !std::isfinite(x[i]) || !std::isfinite(y[i]))
            return fail(nullptr, DYS_ERR_RANGE, "%s: agent %lld has a non-finite age, income or home coordinate", who, (long long)i);
    int rc = pick_device(device, who);
    if (rc != DYS_OK) return rc;
    // normalisers: max |difference| of a column = max - min (exactly: rounded subtraction is monotone); the largest home
    // distance is searched over the DISTINCT home coordinates
    SimParams q{};
    q.n = n;
    const auto mm_a = std::minmax_element(age, age + n), mm_i = std::minmax_element(income, income + n);
    q.max_age = *mm_a.second - *mm_a.first;
    q.max_inc = *mm_i.second - *mm_i.first;
    std::vector<std::pair<double, double>> pts((size_t)n);
    for (int64_t i = 0; i < n; ++i) pts[(size_t)i] = {x[i], y[i]};
    std::sort(pts.begin(), pts.end());
    pts.erase(std::unique(pts.begin(), pts.end()), pts.end());
    const int64_t U = (int64_t)pts.size();
    std::vector<double> ux((size_t)U), uy((size_t)U);
    for (int64_t i = 0; i < U; ++i) { ux[(size_t)i] = pts[(size_t)i].first; uy[(size_t)i] = pts[(size_t)i].second; }
    DevTmp t;
    const double* d_ux = t.up(ux.data(), (size_t)U);
    const double* d_uy = t.up(uy.data(), (size_t)U);
    unsigned long long* d_max = t.get<unsigned long long>(1);
    if (t.e == cudaSuccess) t.e = cudaMemset(d_max, 0, 8);
    unsigned long long bits = 0;
    if (t.e == cudaSuccess) {
        k_max_d2<<<grid_for(U, NET_BLOCK), NET_BLOCK>>>(d_ux, d_uy, U, d_max);
        t.e = cudaMemcpy(&bits, d_max, 8, cudaMemcpyDeviceToHost);
    }
    if (t.e != cudaSuccess) return fail(nullptr, DYS_ERR_CUDA, "%s: %s", who, cudaGetErrorString(t.e));
    double max_d2;
    memcpy(&max_d2, &bits, 8);
    q.max_dist = std::sqrt(max_d2);
    q.w_a = w_a; q.w_i = w_i; q.w_d = w_d; q.min_prob = min_prob;
    const auto unit = [](double w) { return w > 0.0 && w <= 1.0; };
    q.prefilter = unit(w_a) && unit(w_i) && unit(w_d) && min_prob >= 0.0 && min_prob < 1.0 && q.max_age > 0.0 && q.max_inc > 0.0 && q.max_dist > 0.0;
    const double slack = (1.0 - min_prob) * 1.001 + 1e-9;
    q.pre_a = slack * q.max_age;
    q.pre_i = slack * q.max_inc;
    q.pre_d2 = slack * q.max_dist * slack * q.max_dist * 1.001;
    if (maxes) { maxes[0] = q.max_age; maxes[1] = q.max_inc; maxes[2] = q.max_dist; }
    const double* d_age = t.up(age, (size_t)n);
    const double* d_inc = t.up(income, (size_t)n);
    const double* d_x = t.up(x, (size_t)n);
    const double* d_y = t.up(y, (size_t)n);
    const int64_t* d_hh = t.up(household, (size_t)n);
    int64_t* d_cnt = t.get<int64_t>((size_t)n + 1);
    const unsigned grid = grid_for(n, NET_ROWS);
    if (t.e == cudaSuccess) {
        k_sim_edges<false><<<grid, NET_BLOCK>>>(q, d_age, d_inc, d_x, d_y, d_hh, d_cnt, nullptr, nullptr, nullptr);
        t.e = cudaMemcpy(indptr + 1, d_cnt, 8 * (size_t)n, cudaMemcpyDeviceToHost);
    }
    if (t.e != cudaSuccess) return fail(nullptr, DYS_ERR_CUDA, "%s: %s", who, cudaGetErrorString(t.e));
    indptr[0] = 0;
    for (int64_t i = 0; i < n; ++i) indptr[i + 1] += indptr[i];
    const int64_t E = indptr[n];
    if (!col && !prob) return DYS_OK;                                   // sizing call: indptr[n] = number of edges
    if (!col || !prob) return fail(nullptr, DYS_ERR_INVALID_ARG, "%s: col and prob go together", who);
    if (capacity < E) return fail(nullptr, DYS_ERR_RANGE, "%s: %lld edges do not fit the capacity of %lld", who, (long long)E, (long long)capacity);
    if (t.e == cudaSuccess) t.e = cudaMemcpy(d_cnt, indptr, 8 * (size_t)(n + 1), cudaMemcpyHostToDevice);
    int32_t* d_col = t.get<int32_t>((size_t)E);
    double* d_prob = t.get<double>((size_t)E);
    if (t.e == cudaSuccess) {
        k_sim_edges<true><<<grid, NET_BLOCK>>>(q, d_age, d_inc, d_x, d_y, d_hh, nullptr, d_cnt, d_col, d_prob);
        t.e = cudaGetLastError();
        if (t.e == cudaSuccess && E) t.e = cudaMemcpy(col, d_col, 4 * (size_t)E, cudaMemcpyDeviceToHost);
        if (t.e == cudaSuccess && E) t.e = cudaMemcpy(prob, d_prob, 8 * (size_t)E, cudaMemcpyDeviceToHost);
    }
    if (t.e != cudaSuccess) return fail(nullptr, DYS_ERR_CUDA, "%s: %s", who, cudaGetErrorString(t.e));
    return DYS_OK;
}

// numpy's pairwise_sum(a, n) flattened into a stack program (see k_bld_rowsum)
static void pairwise_program(int64_t j0, int64_t n, std::vector<SumOp>& prog)
{
    if (n <= 128) { prog.push_back(SumOp{(int32_t)j0, (int32_t)n}); return; }
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    pairwise_program(j0, n2, prog);
    pairwise_program(j0 + n2, n - n2, prog);
    prog.push_back(SumOp{0, 0});
}

extern "C" int dys_building_edges(int device, int64_t n_bld, const double* floor_space, const double* x, const double* y, double min_prob,
                                  int64_t* indptr /* [n_bld+1] out */, int64_t capacity, int32_t* col, double* prob)
{
    const char* who = "dys_building_edges";
    if (n_bld <= 0 || !floor_space || !x || !y || !indptr) return fail(nullptr, DYS_ERR_INVALID_ARG, "%s: null or empty argument", who);
    if (n_bld > 0x7fffffffll) return fail(nullptr, DYS_ERR_RANGE, "%s: more than 2^31 - 1 buildings", who);
    for (int64_t i = 0; i < n_bld; ++i)
        if (!std::isfinite(floor_space[i]) || !std::isfinite(x[i]) || !std::isfinite(y[i]))
            return fail(nullptr, DYS_ERR_RANGE, "%s: building %lld has a non-finite floor space or coordinate", who, (long long)i);
    int rc = pick_device(device, who);
    if (rc != DYS_OK) return rc;
    std::vector<SumOp> prog;
    pairwise_program(0, n_bld, prog);
    DevTmp t;
    const double* d_fs = t.up(floor_space, (size_t)n_bld);
    const double* d_x = t.up(x, (size_t)n_bld);
    const double* d_y = t.up(y, (size_t)n_bld);
    const SumOp* d_prog = t.up(prog.data(), prog.size());
    double* d_sum = t.get<double>((size_t)n_bld);
    int64_t* d_cnt = t.get<int64_t>((size_t)n_bld + 1);
    const unsigned grid = grid_for(n_bld, NET_ROWS);
    if (t.e == cudaSuccess) {
        k_bld_rowsum<<<grid_for(n_bld, NET_BLOCK), NET_BLOCK>>>(n_bld, d_fs, d_x, d_y, d_prog, (int)prog.size(), d_sum);
        k_bld_edges<false><<<grid, NET_BLOCK>>>(n_bld, min_prob, d_fs, d_x, d_y, d_sum, d_cnt, nullptr, nullptr, nullptr);
        t.e = cudaMemcpy(indptr + 1, d_cnt, 8 * (size_t)n_bld, cudaMemcpyDeviceToHost);
    }
    if (t.e != cudaSuccess) return fail(nullptr, DYS_ERR_CUDA, "%s: %s", who, cudaGetErrorString(t.e));
    indptr[0] = 0;
    for (int64_t i = 0; i < n_bld; ++i) indptr[i + 1] += indptr[i];
    const int64_t E = indptr[n_bld];
    if (!col && !prob) return DYS_OK;                                   // sizing call
    if (!col || !prob) return fail(nullptr, DYS_ERR_INVALID_ARG, "%s: col and prob go together", who);
    if (capacity < E) return fail(nullptr, DYS_ERR_RANGE, "%s: %lld edges do not fit the capacity of %lld", who, (long long)E, (long long)capacity);
    if (t.e == cudaSuccess) t.e = cudaMemcpy(d_cnt, indptr, 8 * (size_t)(n_bld + 1), cudaMemcpyHostToDevice);
    int32_t* d_col = t.get<int32_t>((size_t)E);
    double* d_prob = t.get<double>((size_t)E);
    if (t.e == cudaSuccess) {
        k_bld_edges<true><<<grid, NET_BLOCK>>>(n_bld, min_prob, d_fs, d_x, d_y, d_sum, nullptr, d_cnt, d_col, d_prob);
        t.e = cudaGetLastError();
        if (t.e == cudaSuccess && E) t.e = cudaMemcpy(col, d_col, 4 * (size_t)E, cudaMemcpyDeviceToHost);
        if (t.e == cudaSuccess && E) t.e = cudaMemcpy(prob, d_prob, 8 * (size_t)E, cudaMemcpyDeviceToHost);
    }
    if (t.e != cudaSuccess) return fail(nullptr, DYS_ERR_CUDA, "%s: %s", who, cudaGetErrorString(t.e));
    return DYS_OK;
}

namespace {
struct SpCsrOut {                    // the reachable vertices of every source as CSR instead of dense rows
    int64_t* ptr;                    // [n_sources + 1]
    int64_t capacity;
    int32_t* col;
    double* dist;
    int32_t skip_self;
    double max_dist;                 // keep d < max_dist (+inf: every reachable vertex)
    const uint8_t* col_keep;         // [n] or null: only these vertices are listed
};
}  // namespace

static int shortest_paths_impl(const char* who, int device, int64_t n, const int64_t* indptr, const int32_t* indices, const double* weights,
                               int64_t n_sources, const int32_t* sources, double* out /* [n_sources][n] or null */,
                               const SpCsrOut* csr /* or null */, double* info /* [4], nullable */)
{
    const auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec; };
    const double t_start = now();
    double t_relax = 0.0, t_out = 0.0;
    if (n <= 0 || !indptr || n_sources < 0 || (n_sources && (!sources || (!out && !csr)))) return fail(nullptr, DYS_ERR_INVALID_ARG, "%s: null or empty argument", who);
    if (csr && (!csr->ptr || csr->capacity < 0 || (csr->capacity && (!csr->col || !csr->dist)))) return fail(nullptr, DYS_ERR_INVALID_ARG, "%s: bad output arrays", who);
    if (csr) csr->ptr[0] = 0;
    if (n > 0x7fffffffll) return fail(nullptr, DYS_ERR_RANGE, "%s: more than 2^31 - 1 vertices", who);
    const int64_t E = indptr[n];
    if (indptr[0] != 0 || E < 0 || (E && (!indices || !weights))) return fail(nullptr, DYS_ERR_INVALID_ARG, "%s: bad CSR arrays", who);
    for (int64_t v = 0; v < n; ++v)
        if (indptr[v + 1] < indptr[v]) return fail(nullptr, DYS_ERR_RANGE, "%s: indptr is not monotone at row %lld", who, (long long)v);
    for (int64_t e = 0; e < E; ++e) {
        if (indices[e] < 0 || indices[e] >= n) return fail(nullptr, DYS_ERR_RANGE, "%s: edge %lld points outside the graph", who, (long long)e);
        if (!(weights[e] >= 0.0)) return fail(nullptr, DYS_ERR_RANGE, "%s: edge %lld has a negative or NaN weight", who, (long long)e);
    }
    for (int64_t k = 0; k < n_sources; ++k)
        if (sources[k] < 0 || sources[k] >= n) return fail(nullptr, DYS_ERR_RANGE, "%s: source %lld outside the graph", who, (long long)k);
    if (info) info[0] = info[1] = info[2] = info[3] = 0.0;
    if (n_sources == 0) return DYS_OK;
    int rc = pick_device(device, who);
    if (rc != DYS_OK) return rc;
    // in-edge lists (a vertex pulls from its predecessors): counting sort of the edges by destination, sources ascending
    std::vector<int64_t> in_ptr((size_t)n + 1, 0);
    for (int64_t e = 0; e < E; ++e) ++in_ptr[(size_t)indices[e] + 1];
    for (int64_t v = 0; v < n; ++v) in_ptr[(size_t)v + 1] += in_ptr[(size_t)v];
    std::vector<int32_t> in_src((size_t)E);
    std::vector<double> in_w((size_t)E);
    {
        std::vector<int64_t> cur(in_ptr.begin(), in_ptr.end() - 1);
        for (int64_t u = 0; u < n; ++u)
            for (int64_t e = indptr[u]; e < indptr[u + 1]; ++e) {
                const int64_t at = cur[(size_t)indices[e]]++;
                in_src[(size_t)at] = (int32_t)u;
                in_w[(size_t)at] = weights[e];
            }
    }
    // sources per batch: a multiple of 128 (a warp of k_sp_relax), at most 1024, and D[n][batch] plus its transposed copy within ~4 GB
    int64_t max_batch = 1024;
    if (const char* e = getenv("DYS_SP_BATCH")) max_batch = std::max<int64_t>(SP_CHUNK, std::min<int64_t>(1024, atoll(e) / SP_CHUNK * SP_CHUNK));
    int64_t batch = std::min<int64_t>(max_batch, ((n_sources + SP_CHUNK - 1) / SP_CHUNK) * SP_CHUNK);
    while (batch > SP_CHUNK && (double)n * (double)batch * 16.0 > 4.0e9) batch -= SP_CHUNK;
    const char* e_pv = getenv("DYS_SP_VERTEX_MARKS");                   // developer A/B: one "changed" mark per vertex
    const int fpv = (e_pv && atoi(e_pv)) ? 1 : (int)(batch / SP_CHUNK);
    DevTmp t;
    const int64_t* d_ptr = t.up(in_ptr.data(), (size_t)n + 1);
    const int32_t* d_src = t.up(in_src.data(), (size_t)E);
    const double* d_w = t.up(in_w.data(), (size_t)E);
    double* d_D = t.get<double>((size_t)n * (size_t)batch);
    double* d_out = t.get<double>((size_t)n * (size_t)batch);
    int32_t* d_sources = t.get<int32_t>((size_t)batch);
    int64_t* d_cnt = csr ? t.get<int64_t>((size_t)batch + 1) : nullptr;
    const double bound = csr ? csr->max_dist : (double)INFINITY;
    const uint8_t* d_keep = (csr && csr->col_keep) ? t.up(csr->col_keep, (size_t)n) : nullptr;
    int32_t* d_ccol = nullptr;
    double* d_cdist = nullptr;
    size_t c_room = 0;                                                  // (grown to the largest batch seen)
    std::vector<int64_t> h_cnt((size_t)batch + 1);
    const size_t n_marks = (size_t)n * (size_t)fpv;
    uint8_t* d_dirty = t.get<uint8_t>(2 * n_marks);
    constexpr int GROUP = 8;                                            // sweeps queued between two looks at the `changed` words
    int32_t* d_changed = t.get<int32_t>(GROUP);
    if (t.e != cudaSuccess) return fail(nullptr, DYS_ERR_CUDA, "%s: %s", who, cudaGetErrorString(t.e));
    const unsigned init_grid = grid_for(n * batch, NET_BLOCK);
    const unsigned relax_grid = grid_for(n * (batch / SP_CHUNK) * 32, NET_BLOCK);
    int total_sweeps = 0;
    for (int64_t k0 = 0; k0 < n_sources && t.e == cudaSuccess; k0 += batch) {
        const int nb = (int)std::min<int64_t>(batch, n_sources - k0);
        t.e = cudaMemcpy(d_sources, sources + k0, 4 * (size_t)nb, cudaMemcpyHostToDevice);
        if (t.e == cudaSuccess) t.e = cudaMemset(d_dirty, 0, 2 * n_marks);
        if (t.e != cudaSuccess) break;
        const double t0 = now();
        k_sp_init<<<init_grid, NET_BLOCK>>>(d_D, n, (int)batch, d_sources, nb, d_dirty, fpv);
        int cur = 0;
        bool done = false;
        // every sweep moves the frontier at least one edge further along each shortest path: n sweeps always suffice
        for (int64_t guard = 0; !done && guard <= n + GROUP; guard += GROUP) {
            t.e = cudaMemsetAsync(d_changed, 0, 4 * GROUP);
            for (int g = 0; g < GROUP && t.e == cudaSuccess; ++g) {
                uint8_t* prev = d_dirty + (size_t)cur * n_marks;
                uint8_t* next = d_dirty + (size_t)(cur ^ 1) * n_marks;
                t.e = cudaMemsetAsync(next, 0, n_marks);
                k_sp_relax<<<relax_grid, NET_BLOCK>>>(n, (int)batch, d_ptr, d_src, d_w, d_D, prev, next, d_changed + g, fpv, bound);
                cur ^= 1;
            }
            int32_t ch[GROUP];
            if (t.e == cudaSuccess) t.e = cudaMemcpy(ch, d_changed, sizeof(ch), cudaMemcpyDeviceToHost);
            if (t.e != cudaSuccess) break;
            for (int g = 0; g < GROUP; ++g) {
                if (!ch[g]) { done = true; break; }
                ++total_sweeps;
            }
        }
        if (t.e != cudaSuccess) break;
        if (!done) return fail(nullptr, DYS_ERR_STATE, "%s: the relaxation did not settle", who);
        const double t1 = now();
        t_relax += t1 - t0;
        k_sp_transpose<<<dim3(grid_for(n, 32), (unsigned)(batch / 32)), NET_BLOCK>>>(d_D, n, (int)batch, nb, d_out);
        if (out) t.e = cudaMemcpy(out + k0 * n, d_out, 8 * (size_t)nb * (size_t)n, cudaMemcpyDeviceToHost);
        if (csr && t.e == cudaSuccess) {
            const unsigned cgrid = grid_for((int64_t)nb * 32, NET_BLOCK);
            k_sp_compact<false><<<cgrid, NET_BLOCK>>>(d_out, n, nb, d_sources, csr->skip_self, bound, d_keep, d_cnt + 1, nullptr, nullptr, nullptr);
            t.e = cudaMemcpy(h_cnt.data() + 1, d_cnt + 1, 8 * (size_t)nb, cudaMemcpyDeviceToHost);
            if (t.e != cudaSuccess) break;
            h_cnt[0] = 0;
            for (int k = 0; k < nb; ++k) h_cnt[(size_t)k + 1] += h_cnt[(size_t)k];
            const int64_t m = h_cnt[(size_t)nb], base = csr->ptr[k0];
            if (base + m > csr->capacity)
                return fail(nullptr, DYS_ERR_RANGE, "%s: the reachable pairs do not fit the capacity of %lld (at least %lld after %lld sources)", who,
                            (long long)csr->capacity, (long long)(base + m), (long long)(k0 + nb));
            for (int k = 0; k < nb; ++k) csr->ptr[k0 + k + 1] = base + h_cnt[(size_t)k + 1];
            if (m > 0) {
                if ((size_t)m > c_room) {
                    c_room = (size_t)m + (size_t)m / 4;
                    d_ccol = t.get<int32_t>(c_room);
                    d_cdist = t.get<double>(c_room);
                }
                if (t.e == cudaSuccess) t.e = cudaMemcpy(d_cnt, h_cnt.data(), 8 * (size_t)(nb + 1), cudaMemcpyHostToDevice);
                if (t.e != cudaSuccess) break;
                k_sp_compact<true><<<cgrid, NET_BLOCK>>>(d_out, n, nb, d_sources, csr->skip_self, bound, d_keep, nullptr, d_cnt, d_ccol, d_cdist);
                t.e = cudaMemcpy(csr->col + base, d_ccol, 4 * (size_t)m, cudaMemcpyDeviceToHost);
                if (t.e == cudaSuccess) t.e = cudaMemcpy(csr->dist + base, d_cdist, 8 * (size_t)m, cudaMemcpyDeviceToHost);
            }
        }
        t_out += now() - t1;
    }
    if (t.e != cudaSuccess) return fail(nullptr, DYS_ERR_CUDA, "%s: %s", who, cudaGetErrorString(t.e));
    if (info) { info[0] = (double)total_sweeps; info[1] = now() - t_start - t_relax - t_out; info[2] = t_relax; info[3] = t_out; }
    return DYS_OK;
}

extern "C" int dys_shortest_paths(int device, int64_t n, const int64_t* indptr, const int32_t* indices, const double* weights,
                                  int64_t n_sources, const int32_t* sources, double* out /* [n_sources][n] */, double* info /* [4], nullable */)
{
    if (n_sources > 0 && !out) return fail(nullptr, DYS_ERR_INVALID_ARG, "dys_shortest_paths: null out");
    return shortest_paths_impl("dys_shortest_paths", device, n, indptr, indices, weights, n_sources, sources, out, nullptr, info);
}

extern "C" int dys_shortest_paths_csr(int device, int64_t n, const int64_t* indptr, const int32_t* indices, const double* weights,
                                      int64_t n_sources, const int32_t* sources, int32_t skip_self, double max_dist,
                                      const uint8_t* col_keep /* [n], nullable */, int64_t* out_ptr /* [n_sources + 1] */,
                                      int64_t capacity, int32_t* out_col, double* out_dist, double* info /* [4], nullable */)
{
    if (!out_ptr) return fail(nullptr, DYS_ERR_INVALID_ARG, "dys_shortest_paths_csr: null out_ptr");
    if (!(max_dist > 0.0)) return fail(nullptr, DYS_ERR_INVALID_ARG, "dys_shortest_paths_csr: max_dist must be positive (+inf for no bound)");
    const SpCsrOut csr{out_ptr, capacity, out_col, out_dist, skip_self, max_dist, col_keep};
    return shortest_paths_impl("dys_shortest_paths_csr", device, n, indptr, indices, weights, n_sources, sources, nullptr, &csr, info);
}

extern "C" const char* dys_version(void) { return DYS_CHECKS ? "dysturbd-b200 0.2 (sm_100a, checked build)" : "dysturbd-b200 0.2 (sm_100a)"; }

extern "C" const char* dys_last_error(const dys_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" int dys_create(const dys_params* params, int device, dys_handle** out)
{
    if (!params || !out) return fail(nullptr, DYS_ERR_INVALID_ARG, "dys_create: null argument");
    const dys_params& p = *params;
    if (p.n_agents <= 0 || p.n_agents > (int64_t)COL_MASK + 1) return fail(nullptr, DYS_ERR_INVALID_ARG, "n_agents must be in [1, 2^27]");
    if (p.n_households <= 0 || p.n_buildings <= 0 || p.n_areas < 0) return fail(nullptr, DYS_ERR_INVALID_ARG, "bad H/B/S");
    if (p.n_slots < 1 || p.n_slots > 16) return fail(nullptr, DYS_ERR_INVALID_ARG, "n_slots must be in [1,16]");
    // the day-ahead household prediction needs diagnosis >= 1; compute_R (contagious3.py:29-42) sums bins 1..recover-1 of
    // the 21-bin days-since-infection histogram
    if (p.diagnosis < 1 || p.quarantine < 1 || p.recover < 1 || p.recover > HIST_LEN || p.hospital_recover < 1)
        return fail(nullptr, DYS_ERR_INVALID_ARG, "need diagnosis >= 1, quarantine >= 1, 1 <= recover <= %d, hospital_recover >= 1", HIST_LEN);
    if (!(p.compliance >= 0.0 && p.compliance <= 1.0)) return fail(nullptr, DYS_ERR_INVALID_ARG, "compliance must be in [0, 1]");
    const int world = p.world > 1 ? p.world : 1;
    const int R = p.n_replicas > 1 ? p.n_replicas : 1;
    if (R > 4096) return fail(nullptr, DYS_ERR_INVALID_ARG, "n_replicas must be <= 4096");
    if (R > 1 && (world > 1 || (p.flags & (DYS_WANT_IO_MAT | DYS_TIME_KERNELS))))
        return fail(nullptr, DYS_ERR_INVALID_ARG, "replicas are single-GPU (spread handles over GPUs) and exclude DYS_WANT_IO_MAT / DYS_TIME_KERNELS");
    int64_t lo = 0, hi = p.n_agents;
    if (world > 1) {
        lo = p.shard_lo; hi = p.shard_hi;
        if (lo < 0 || hi > p.n_agents || lo >= hi || (lo % 1024) != 0) return fail(nullptr, DYS_ERR_INVALID_ARG, "bad shard range (lo must be a multiple of 1024)");
        if (p.flags & DYS_WANT_IO_MAT) return fail(nullptr, DYS_ERR_INVALID_ARG, "DYS_WANT_IO_MAT is not available in sharded mode");
    }
    if ((p.flags & DYS_WANT_IO_MAT) && p.n_areas * p.n_areas > (int64_t)1 << 26)
        return fail(nullptr, DYS_ERR_INVALID_ARG, "DYS_WANT_IO_MAT needs S*S <= 2^26; use dys_get_events");
    if (p.halo_hi > p.halo_lo && (p.halo_lo < 0 || p.halo_hi > p.n_agents || p.halo_lo > lo || p.halo_hi < hi))
        return fail(nullptr, DYS_ERR_INVALID_ARG, "halo must contain the owned range");
    if (p.hh_hi > p.n_households || p.bld_hi > p.n_buildings || p.area_hi > p.n_areas || p.hh_lo < 0 || p.bld_lo < 0 || p.area_lo < 0)
        return fail(nullptr, DYS_ERR_INVALID_ARG, "owned household / building / area range outside the population");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, DYS_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(nullptr, DYS_ERR_INVALID_ARG, "device %d out of range (%d devices)", device, ndev);
    dys_handle* h = new dys_handle();      // (from here on every failure goes through dys_destroy)
    h->p = p;
    h->device = device;
    h->R = R;
    for (int i = 0; i < DYS_OUT_RING; ++i) { h->out_day[i] = -1; h->out_step[i] = -1; h->out_ev[i] = nullptr; h->out_ev_of[i] = i; }
    for (int i = 0; i < DYS_STATS_RING; ++i) h->hist_day[i] = -1;
    h->hist_stats.assign((size_t)DYS_STATS_RING * R * N_STATS, 0);
    h->rep.assign((size_t)R, RepParam{p.norm_factor, p.compliance, p.seed});
    { const char* nf = getenv("DYS_NO_FUSE"); h->no_fuse = nf && nf[0] == '1'; }
    h->trace_path = getenv("DYS_TRACE");
    h->test_hooks = getenv("DYS_TEST_HANG_BLOCK") != nullptr;
    { const char* t = getenv("DYS_UPLOAD_ALWAYS"); h->upload_always = t && t[0] == '1'; }
    { const char* t = getenv("DYS_PLAN_MAX"); if (t) h->plan_max = atoll(t); }
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->up_stream, cudaStreamNonBlocking);
    for (int k = 0; k < DYS_OUT_RING && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&h->day_done[k], cudaEventDisableTiming);
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
        e = cudaEventCreateWithFlags(&h->up_done[k], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ring_free[k], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) { fail(nullptr, DYS_ERR_CUDA, "cuda init: %s", cudaGetErrorString(e)); dys_destroy(h); return DYS_ERR_CUDA; }
    DevCtx& c = h->c;
    c.n_loc = hi - lo; c.n_pad = round_up(c.n_loc, 1024); c.n_glob = p.n_agents; c.lo = lo;
    c.rt_lo = 0; c.rt_n = p.n_agents;
    if (p.halo_hi > p.halo_lo) { c.rt_lo = p.halo_lo; c.rt_n = p.halo_hi - p.halo_lo; }
    c.H = p.n_households; c.B = p.n_buildings; c.S = p.n_areas; c.V = p.n_slots;
    c.B_out = c.B; c.S_out = c.S;
    h->hh_lo = h->bld_lo = h->area_lo = 0;
    if (p.hh_hi > p.hh_lo) { h->hh_lo = p.hh_lo; c.H = p.hh_hi - p.hh_lo; }
    if (p.bld_hi > p.bld_lo) { h->bld_lo = p.bld_lo; c.B_out = p.bld_hi - p.bld_lo; }
    if (p.area_hi > p.area_lo) { h->area_lo = p.area_lo; c.S_out = p.area_hi - p.area_lo; }
    c.recover = p.recover; c.hospital_recover = p.hospital_recover; c.diagnosis = p.diagnosis; c.quarantine = p.quarantine;
    c.adm_min = p.adm_min_day; c.adm_max = p.adm_max_day; c.death_min = p.death_min_adm_day;
    c.norm_factor = p.norm_factor; c.compliance = p.compliance; c.seed = p.seed;
    c.n_rep = R; c.rep = 0; c.r_idle_day = 0;
    {
        const char* t = getenv("DYS_SPIN_TIMEOUT_MS");
        const double ms = t ? atof(t) : 20000.0;
        c.spin_timeout_ns = ms > 0 ? (unsigned long long)(ms * 1e6) : 0ull;
    }
    int rc = DYS_OK;
    auto A = [&](int r) { if (rc == DYS_OK) rc = r; };
    double* d_pdf = nullptr;
    // everything mutable exists once per replica, replica r at r * stride (shift_replica)
    A(dev_alloc(h, &c.status, c.n_pad * R)); A(dev_alloc(h, &c.rec, c.n_pad * R)); A(dev_alloc(h, &c.src, c.n_pad * R));
    A(dev_alloc(h, (int32_t**)&c.hh, c.n_pad));
    A(dev_alloc(h, (double**)&c.risk, c.n_pad)); A(dev_alloc(h, (double**)&c.p_adm, c.n_pad)); A(dev_alloc(h, (double**)&c.p_death, c.n_pad));
    A(dev_alloc(h, (int32_t**)&c.routine, c.rt_n * c.V));
    h->ib_stride = round_up(c.n_glob, 1024) + 1024;
    c.ib_stride = h->ib_stride;
    A(dev_alloc(h, &h->d_ib, 2 * h->ib_stride * R));
    A(dev_alloc(h, &h->d_arrive, 8));
    A(dev_alloc(h, (uint8_t**)&c.open, c.B));
    for (int k = 0; k < 2; ++k) {
        A(dev_alloc(h, &h->buf.att[k], (c.n_pad + 16) * R));
        A(dev_alloc(h, &h->buf.mb_any[k], c.n_pad * R)); A(dev_alloc(h, &h->buf.mb_iso[k], c.n_pad * R));
        h->buf.ib[k] = h->d_ib + (size_t)k * h->ib_stride * R;
        for (int r = 0; r < 8; ++r) h->buf.peer_ib[r][k] = h->buf.ib[k];
    }
    for (int k = 0; k < Q_RING; ++k) A(dev_alloc(h, &h->buf.qflag[k], (c.H + 16) * R));
    A(dev_alloc(h, &h->buf.df, (int64_t)DF_RING * DF_LEN * R));
    A(dev_alloc(h, &h->buf.part, (int64_t)PART_RING * N_SLOTS * N_STATS * R)); A(dev_alloc(h, &c.ticket, 1));
    A(dev_alloc(h, (int64_t**)&c.hh_ptr, c.H + 1)); A(dev_alloc(h, (int32_t**)&c.hh_mem, c.n_pad));
    A(dev_alloc(h, &h->d_rep, R));
    c.rep_params = R > 1 ? h->d_rep : nullptr;
    c.n_peers = 1; c.rank = p.rank; c.arrive = h->d_arrive; c.n_wr = 1;
    A(dev_alloc(h, &d_pdf, DYS_PDF_LEN));
    A(dev_alloc(h, &h->d_bad, 4));
    {   // one output block per ring slot and replica: everything a day produces leaves the device in a single copy
        auto up16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
        h->off_sa = sizeof(int64_t) * N_STATS;
        h->off_bc = h->off_sa + ((p.flags & DYS_WANT_SA_HIST) ? up16(sizeof(int32_t) * (size_t)(c.S_out * HIST_LEN)) : 0);
        h->off_ev = h->off_bc + ((p.flags & DYS_WANT_BUILDING_COUNTS) ? up16(sizeof(int32_t) * (size_t)c.B_out) : 0);
        h->ev_inline = (p.flags & DYS_WANT_EVENTS) ? (c.n_pad / 64 > 1024 ? c.n_pad / 64 : 1024) : 0;
        if (h->ev_inline > c.n_pad) h->ev_inline = c.n_pad;
        h->out_copy_bytes = h->off_ev + sizeof(int2) * (size_t)h->ev_inline;
        h->out_dev_bytes = h->off_ev + ((p.flags & DYS_WANT_EVENTS) ? sizeof(int2) * (size_t)c.n_pad : 0);
        c.out_stride = (int64_t)h->out_dev_bytes;
        // (one allocation: the slots of a window are neighbours, so a whole window ships in one or two 2-D copies)
        A(dev_alloc(h, &h->d_out[0], (int64_t)h->out_dev_bytes * R * DYS_OUT_RING));
        for (int k = 1; k < DYS_OUT_RING; ++k) h->d_out[k] = h->d_out[0] ? h->d_out[0] + (size_t)k * h->out_dev_bytes * R : nullptr;
    }
    if (p.flags & DYS_WANT_IO_MAT) A(dev_alloc(h, &c.io_mat, c.S * c.S));
    c.pdf = d_pdf;
    c.n_closed = 0;
    if (rc == DYS_OK) {
        int per_sm = 0, sms = 0;
        cudaError_t e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_push, BLOCK, 0);
        if (e2 == cudaSuccess) e2 = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        if (e2 != cudaSuccess || per_sm < 1) { h->err = std::string("k_push occupancy: ") + cudaGetErrorString(e2); rc = DYS_ERR_CUDA; }
        const int64_t n_cta = (c.rt_n + 15) / 16;                                       // a block's share is at least 16 agents
        const int64_t want = (int64_t)per_sm * sms;                                     // persistent: a multiple of the SM count
        h->push_grid = (int)(n_cta < want ? n_cta : want);
        int coop = 0, per_sm_days = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
        if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_days, k_days<true, true>, 256, 0) == cudaSuccess && per_sm_days > 0)
            h->days_grid = per_sm_days * sms;
        if (R > 1 && h->days_grid < 2) { h->err = "replicas need the cooperative k_days launch"; rc = DYS_ERR_NOT_AVAILABLE; }
    }
    if (rc == DYS_OK && h->days_grid > 1)
        for (int k = 0; k < 3; ++k) {      // the daily equal-work cut of k_days (see plan_cut), over [replica][piece]
            A(dev_alloc(h, &h->buf.bounds[k], h->days_grid + 1));
            if (k < 2) A(dev_alloc(h, &h->buf.piece_work[k], c.n_pad / 128 * R));
        }
    for (int k = 0; k < 2; ++k) {
        A(dev_alloc(h, &h->d_open_ring[k], (int64_t)MAX_FUSED_DAYS * c.B));
        if (rc == DYS_OK && cudaMallocHost((void**)&h->h_open_ring[k], (size_t)MAX_FUSED_DAYS * (size_t)c.B) != cudaSuccess) rc = DYS_ERR_CUDA;
    }
    {   // barrier counters and work words in one allocation: one memset per window
        unsigned long long* blk = nullptr;
        A(dev_alloc(h, &blk, 2 + (h->days_grid > 0 ? h->days_grid : 1)));
        h->d_barrier = (unsigned int*)blk;
        h->d_work = blk ? blk + 2 : nullptr;
    }
    { const char* t = getenv("DYS_NO_STEAL"); h->no_steal = t && t[0] == '1'; }      // (developer: ablation)
    if (rc == DYS_OK) {
        void* ab = nullptr;
        if (cudaHostAlloc(&ab, 64, cudaHostAllocMapped) != cudaSuccess) rc = DYS_ERR_CUDA;
        else {
            h->h_abort = (volatile int32_t*)ab;
            *h->h_abort = 0;
            void* dab = nullptr;
            if (cudaHostGetDevicePointer(&dab, ab, 0) != cudaSuccess) rc = DYS_ERR_CUDA;
            c.abort_host = (volatile int32_t*)dab;
            int32_t* dword = nullptr;
            A(dev_alloc(h, &dword, 16));
            c.abort_word = dword;
        }
    }
    if (rc == DYS_OK && cudaMemcpyAsync(d_pdf, p.pdf_table, sizeof(double) * DYS_PDF_LEN, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) rc = DYS_ERR_CUDA;
    if (rc == DYS_OK && cudaMemcpyAsync(h->d_rep, h->rep.data(), sizeof(RepParam) * (size_t)R, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) rc = DYS_ERR_CUDA;
    if (rc == DYS_OK && cudaMemsetAsync((void*)c.open, 1, (size_t)c.B, h->stream) != cudaSuccess) rc = DYS_ERR_CUDA;
    if (rc == DYS_OK && cudaMallocHost((void**)&h->h_out, h->out_copy_bytes * DYS_OUT_RING * (size_t)R) != cudaSuccess) rc = DYS_ERR_CUDA;
    if (rc == DYS_OK && cudaMallocHost((void**)&h->h_open, (size_t)c.B) != cudaSuccess) rc = DYS_ERR_CUDA;
    if (rc == DYS_OK) memset(h->h_open, 1, (size_t)c.B);
    if (rc == DYS_OK && cudaMallocHost((void**)&h->h_small, sizeof(int32_t) * 16) != cudaSuccess) rc = DYS_ERR_CUDA;
    for (int i = 0; i < DYS_OUT_RING; ++i)
        if (rc == DYS_OK && cudaEventCreateWithFlags(&h->out_ev[i], cudaEventDisableTiming) != cudaSuccess) rc = DYS_ERR_CUDA;
    if (rc == DYS_OK && (p.flags & DYS_TIME_KERNELS)) {
        h->tev.resize(dys_handle::T_RING * 4);
        for (auto& ev : h->tev) if (cudaEventCreate(&ev) != cudaSuccess) { rc = DYS_ERR_CUDA; break; }
    }
    if (rc == DYS_OK && cudaStreamSynchronize(h->stream) != cudaSuccess) rc = DYS_ERR_CUDA;
    if (rc != DYS_OK) {
        fail(nullptr, rc, "dys_create: allocation failed: %s", h->err.empty() ? cudaGetErrorString(cudaGetLastError()) : h->err.c_str());
        dys_destroy(h);
        return rc;
    }
    *out = h;
    return DYS_OK;
}

extern "C" int dys_destroy(dys_handle* h)
{
    if (!h) return DYS_OK;
    cudaSetDevice(h->device);
    if (h->h_abort) dys_abort(h);         // a launch that is still spinning ends
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->up_stream) { cudaStreamSynchronize(h->up_stream); cudaStreamDestroy(h->up_stream); }
    for (void* p : h->allocs) cudaFree(p);
    if (h->h_out) cudaFreeHost(h->h_out);
    if (h->h_open) cudaFreeHost(h->h_open);
    if (h->h_small) cudaFreeHost(h->h_small);
    for (int i = 0; i < DYS_OUT_RING; ++i) if (h->out_ev[i]) cudaEventDestroy(h->out_ev[i]);
    for (auto ev : h->tev) if (ev) cudaEventDestroy(ev);
    if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
    for (int k = 0; k < DYS_OUT_RING; ++k) if (h->day_done[k]) cudaEventDestroy(h->day_done[k]);
    for (int k = 0; k < 2; ++k) {
        if (h->h_open_ring[k]) cudaFreeHost(h->h_open_ring[k]);
        if (h->up_done[k]) cudaEventDestroy(h->up_done[k]);
        if (h->ring_free[k]) cudaEventDestroy(h->ring_free[k]);
    }
    if (h->h_abort) cudaFreeHost((void*)h->h_abort);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return DYS_OK;
}

// device lockdown policy at the start of a run: day 0 runs under the loaded flags, days 1 and 2 are decided from Known R = 0
// (all open, contagious3.py:57-60 with vR = 0); no R history yet
static int init_lockdown(dys_handle* h)
{
    DevCtx& c = h->c;
    const size_t R = (size_t)h->R, B = (size_t)c.B;
    CK(h, cudaMemsetAsync(h->ld.flags, 1, 4 * R * B, h->stream));
    for (size_t r = 0; r < R; ++r) CK(h, cudaMemcpyAsync(h->ld.flags + r * B, c.open, B, cudaMemcpyDeviceToDevice, h->stream));
    std::vector<int32_t> nc(4 * R, 0);
    for (size_t r = 0; r < R; ++r) nc[r] = c.n_closed;
    CK(h, cudaMemcpyAsync(h->ld.n_closed, nc.data(), sizeof(int32_t) * nc.size(), cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaMemsetAsync(h->ld.r_hist, 0, sizeof(double) * 16 * R, h->stream));
    if (h->ld.r_sa) CK(h, cudaMemsetAsync(h->ld.r_sa, 0, sizeof(double) * 16 * R * (size_t)h->ld.S, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return DYS_OK;
}

// peer mode with a boundary split runs two service blocks (fold / plan and halo push), see k_days
static int n_service_blocks(const dys_handle* h) { return (DYS_BOUNDARY_FIRST && h->n_peers > 1 && h->c.bnd_ok && h->days_grid > 4 * HALO_BLOCKS) ? 1 + HALO_BLOCKS : 1; }

// until measured costs exist the work blocks of k_days get equal runs of pieces
static int init_cut(dys_handle* h)
{
    DevCtx& c = h->c;
    if (!h->buf.bounds[0]) return DYS_OK;
    const int n_work = h->days_grid - n_service_blocks(h);
    const int64_t n_pieces = c.n_pad / 128 * (int64_t)h->R, per = (n_pieces + n_work - 1) / n_work;
    std::vector<int32_t> cut((size_t)h->days_grid + 1, (int32_t)n_pieces);
    for (int b = 0; b <= n_work; ++b) cut[(size_t)b] = (int32_t)((int64_t)b * per < n_pieces ? (int64_t)b * per : n_pieces);
    for (int k = 0; k < 3; ++k) {
        CK(h, cudaMemcpyAsync(h->buf.bounds[k], cut.data(), sizeof(int32_t) * cut.size(), cudaMemcpyHostToDevice, h->stream));
        if (k < 2) CK(h, cudaMemsetAsync(h->buf.piece_work[k], 0, sizeof(uint32_t) * (size_t)n_pieces, h->stream));
    }
    CK(h, cudaStreamSynchronize(h->stream));      // `cut` leaves scope
    return DYS_OK;
}

// state was (re)loaded: the snapshot, attention bytes, mailboxes, flag rings, partial sums and the cut start clean
static int reset_day_scratch(dys_handle* h)
{
    DevCtx& c = h->c;
    const size_t R = (size_t)h->R;
    for (int k = 0; k < 2; ++k) {
        CK(h, cudaMemsetAsync(h->buf.att[k], 0, ((size_t)c.n_pad + 16) * R, h->stream));
        CK(h, cudaMemsetAsync(h->buf.mb_any[k], 0xFF, sizeof(int32_t) * (size_t)c.n_pad * R, h->stream));   // -1: empty mailbox
        CK(h, cudaMemsetAsync(h->buf.mb_iso[k], 0xFF, sizeof(int32_t) * (size_t)c.n_pad * R, h->stream));
    }
    for (int k = 0; k < Q_RING; ++k) CK(h, cudaMemsetAsync(h->buf.qflag[k], 0, ((size_t)c.H + 16) * R, h->stream));
    CK(h, cudaMemsetAsync(h->buf.df, 0, sizeof(int32_t) * DF_RING * DF_LEN * R, h->stream));
    CK(h, cudaMemsetAsync(h->buf.part, 0, sizeof(unsigned long long) * PART_RING * N_SLOTS * N_STATS * R, h->stream));
    CK(h, cudaMemsetAsync(c.ticket, 0, sizeof(unsigned int), h->stream));
    { int rcut = init_cut(h); if (rcut != DYS_OK) return rcut; }
    if (h->ld.mode) { int rl = init_lockdown(h); if (rl != DYS_OK) return rl; }
    h->ib_day = -1;
    h->begun = false;
    CK(h, cudaMemsetAsync(h->d_arrive, 0, sizeof(unsigned long long) * 8, h->stream));   // (peer mode: the caller barriers across ranks before stepping)
    h->pub_epoch = 0;
    CK(h, cudaStreamSynchronize(h->stream));
    CK(h, cudaStreamSynchronize(h->copy_stream));
    CK(h, cudaStreamSynchronize(h->up_stream));
    for (int i = 0; i < DYS_OUT_RING; ++i) { h->out_day[i] = -1; h->out_step[i] = -1; }
    for (int i = 0; i < DYS_STATS_RING; ++i) h->hist_day[i] = -1;
    if (h->h_abort) {
        *h->h_abort = 0;
        CK(h, cudaMemsetAsync((void*)h->c.abort_word, 0, sizeof(int32_t), h->stream));
        CK(h, cudaStreamSynchronize(h->stream));
    }
    return DYS_OK;
}

static int check_bad(dys_handle* h, const char* what)
{
    CK(h, cudaMemcpyAsync(h->h_small, h->d_bad, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    const int bad = h->h_small[0];
    if (bad) {
        CK(h, cudaMemsetAsync(h->d_bad, 0, 2 * sizeof(int32_t), h->stream));
        return fail(h, DYS_ERR_RANGE, "%s: invalid input (flags 0x%x: 1 status code, 2 hh/home/sa index, 4 t_adm != 0 for a "
                    "never-admitted status, 8 src index, 16 routine building index, 32 rowptr not monotone, 64 column index, "
                    "128 columns not strictly ascending, 256 day out of int16 range, 512 susceptible agent with an infector, "
                    "1024 household with more than 32767 members, 2048 trig_ptr not monotone, 4096 trig_hh outside the owned "
                    "households)", what, bad);
    }
    return DYS_OK;
}

static int load_days(dys_handle* h, AgentRec* rec, int field, const int32_t* src, int32_t* tmp, int64_t n)
{
    int rc = copy_in(h, tmp, src, sizeof(int32_t) * (size_t)n);
    if (rc != DYS_OK) return rc;
    k_narrow_days<<<grid_for(n, 256), 256, 0, h->stream>>>(tmp, rec, field, n, h->d_bad);
    CK(h, cudaGetLastError());
    return DYS_OK;
}

// the state columns of ONE replica (the context's pointers are replica 0's)
static int load_state(dys_handle* h, const uint8_t* status_x2, const int32_t* t_inf, const int32_t* t_q, const int32_t* t_adm,
                      const int32_t* src, int r = 0)
{
    DevCtx c = h->c;
    c.status += (size_t)c.n_pad * r; c.rec += (size_t)c.n_pad * r; c.src += (size_t)c.n_pad * r;
    const int64_t n = c.n_loc;
    int32_t* tmp = nullptr;
    CK(h, cudaMalloc((void**)&tmp, sizeof(int32_t) * (size_t)n));
    int rc = DYS_OK;
    auto A = [&](int r) { if (rc == DYS_OK) rc = r; };
    A(copy_in(h, c.status, status_x2, (size_t)n));
    A(load_days(h, c.rec, 0, t_inf, tmp, n));
    A(load_days(h, c.rec, 1, t_q, tmp, n));
    A(load_days(h, c.rec, 2, t_adm, tmp, n));
    A(copy_in(h, c.src, src, sizeof(int32_t) * (size_t)n));
    cudaStreamSynchronize(h->stream);
    cudaFree(tmp);
    return rc;
}

// replica 0 -> every other replica (state and the static half of the agent records)
static int replicate_state(dys_handle* h)
{
    DevCtx& c = h->c;
    for (int r = 1; r < h->R; ++r) {
        CK(h, cudaMemcpyAsync(c.status + (size_t)c.n_pad * r, c.status, (size_t)c.n_pad, cudaMemcpyDeviceToDevice, h->stream));
        CK(h, cudaMemcpyAsync(c.rec + (size_t)c.n_pad * r, c.rec, sizeof(AgentRec) * (size_t)c.n_pad, cudaMemcpyDeviceToDevice, h->stream));
        CK(h, cudaMemcpyAsync(c.src + (size_t)c.n_pad * r, c.src, sizeof(int32_t) * (size_t)c.n_pad, cudaMemcpyDeviceToDevice, h->stream));
    }
    return DYS_OK;
}

// validation of the loaded state; also when every loaded recovered agent has left the days-since-infection histogram
static int validate_state(dys_handle* h, const char* what, int r = 0)
{
    DevCtx c = h->c;
    c.status += (size_t)c.n_pad * r; c.rec += (size_t)c.n_pad * r; c.src += (size_t)c.n_pad * r;
    k_validate_population<<<grid_for(c.n_loc, 256), 256, 0, h->stream>>>(c, h->d_bad);
    CK(h, cudaGetLastError());
    int rc = check_bad(h, what);
    if (rc != DYS_OK) return rc;
    if (h->h_small[1] > h->c.r_idle_day) h->c.r_idle_day = h->h_small[1];
    CK(h, cudaMemsetAsync(h->d_bad, 0, 2 * sizeof(int32_t), h->stream));
    return DYS_OK;
}

extern "C" int dys_load_population(dys_handle* h, const uint8_t* status_x2, const int32_t* t_inf, const int32_t* t_q,
                                   const int32_t* t_adm, const int32_t* src, const int32_t* hh, const int32_t* home,
                                   const int32_t* sa, const double* risk, const double* p_adm, const double* p_death,
                                   const int32_t* routine)
{
    if (!h) return DYS_ERR_INVALID_ARG;
    if (!status_x2 || !t_inf || !t_q || !t_adm || !src || !hh || !home || !sa || !risk || !p_adm || !p_death || !routine)
        return fail(h, DYS_ERR_INVALID_ARG, "dys_load_population: null array");
    CK(h, cudaSetDevice(h->device));
    DevCtx& c = h->c;
    const size_t n = (size_t)c.n_loc;
    int rc = load_state(h, status_x2, t_inf, t_q, t_adm, src);
    auto A = [&](int r) { if (rc == DYS_OK) rc = r; };
    int32_t *d_home = nullptr, *d_sa = nullptr;      // staging: both end up inside the agent records
    if (cudaMalloc((void**)&d_home, 4 * n) != cudaSuccess || cudaMalloc((void**)&d_sa, 4 * n) != cudaSuccess) {
        cudaFree(d_home);
        return fail(h, DYS_ERR_NOMEM, "dys_load_population: staging allocation failed");
    }
    A(copy_in(h, (void*)c.hh, hh, 4 * n)); A(copy_in(h, d_home, home, 4 * n)); A(copy_in(h, d_sa, sa, 4 * n));
    A(copy_in(h, (void*)c.risk, risk, 8 * n)); A(copy_in(h, (void*)c.p_adm, p_adm, 8 * n)); A(copy_in(h, (void*)c.p_death, p_death, 8 * n));
    A(copy_in(h, (void*)c.routine, routine, 4 * (size_t)c.rt_n * c.V));
    if (rc == DYS_OK)
        k_localize<<<grid_for(c.n_loc, 256), 256, 0, h->stream>>>((int32_t*)c.hh, d_home, d_sa, c.rec, c.n_loc, (int32_t)h->hh_lo,
                                                                  (int32_t)h->bld_lo, (int32_t)h->area_lo);
    cudaStreamSynchronize(h->stream);
    cudaFree(d_home); cudaFree(d_sa);
    if (rc != DYS_OK) return rc;
    k_validate_routine<<<grid_for(c.rt_n * c.V, 256), 256, 0, h->stream>>>(c.routine, c.rt_n * c.V, c.B, h->d_bad);
    CK(h, cudaGetLastError());
    c.r_idle_day = 0;
    rc = validate_state(h, "dys_load_population");
    if (rc != DYS_OK) return rc;
    {   // household -> members (phase F flags the MEMBERS of a diagnosed agent's household, contagious3.py:263)
        int32_t* d_cnt = nullptr;
        unsigned long long* d_cursor = nullptr;
        cudaError_t e = cudaMalloc((void**)&d_cnt, sizeof(int32_t) * (size_t)c.H);
        if (e == cudaSuccess) e = cudaMalloc((void**)&d_cursor, sizeof(unsigned long long) * (size_t)c.H);
        if (e == cudaSuccess) e = cudaMemsetAsync(d_cnt, 0, sizeof(int32_t) * (size_t)c.H, h->stream);
        if (e == cudaSuccess) {
            k_count_members<<<grid_for(c.n_loc, 256), 256, 0, h->stream>>>(c.hh, c.n_loc, d_cnt);
            k_scan_counts<<<1, 1024, 0, h->stream>>>(d_cnt, c.H, (int64_t*)c.hh_ptr, d_cursor);
            k_fill_members<<<grid_for(c.n_loc, 256), 256, 0, h->stream>>>(c.hh, c.n_loc, d_cursor, (int32_t*)c.hh_mem);
            k_pack_records<<<grid_for(c.n_loc, 256), 256, 0, h->stream>>>(c.hh, c.hh_ptr, c.p_adm, c.rec, c.n_loc, h->d_bad);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        cudaFree(d_cnt); cudaFree(d_cursor);
        if (e != cudaSuccess) return fail(h, DYS_ERR_CUDA, "building the household index failed: %s", cudaGetErrorString(e));
        rc = check_bad(h, "dys_load_population");
        if (rc != DYS_OK) return rc;
    }
    h->have_pop = true;
    h->have_graph = false;
    rc = replicate_state(h);
    return rc == DYS_OK ? reset_day_scratch(h) : rc;
}

extern "C" int dys_set_state(dys_handle* h, const uint8_t* status_x2, const int32_t* t_inf, const int32_t* t_q,
                             const int32_t* t_adm, const int32_t* src)
{
    if (!h) return DYS_ERR_INVALID_ARG;
    if (!h->have_pop) return fail(h, DYS_ERR_STATE, "dys_set_state before dys_load_population");
    if (!status_x2 || !t_inf || !t_q || !t_adm || !src) return fail(h, DYS_ERR_INVALID_ARG, "dys_set_state: null array");
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaStreamSynchronize(h->stream));
    const int r = h->sel >= 0 ? h->sel : 0;
    if (h->sel < 0) h->c.r_idle_day = 0;      // (a single replica's reload can only push the day later)
    int rc = load_state(h, status_x2, t_inf, t_q, t_adm, src, r);
    if (rc != DYS_OK) return rc;
    rc = validate_state(h, "dys_set_state", r);
    if (rc != DYS_OK) return rc;
    if (h->sel < 0) { rc = replicate_state(h); if (rc != DYS_OK) return rc; }
    h->last_day = -1;
    return reset_day_scratch(h);
}

extern "C" int dys_set_run(dys_handle* h, double norm_factor, double compliance, uint64_t seed)
{
    if (!h) return DYS_ERR_INVALID_ARG;
    if (!(compliance >= 0.0 && compliance <= 1.0)) return fail(h, DYS_ERR_INVALID_ARG, "compliance must be in [0, 1]");
    h->c.norm_factor = norm_factor; h->c.compliance = compliance; h->c.seed = seed;
    h->p.norm_factor = norm_factor; h->p.compliance = compliance; h->p.seed = seed;
    if (h->R == 1) h->rep[0] = RepParam{norm_factor, compliance, seed};
    return DYS_OK;
}

extern "C" int dys_set_replica_params(dys_handle* h, const double* norm_factor, const double* compliance, const uint64_t* seed)
{
    if (!h) return DYS_ERR_INVALID_ARG;
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaStreamSynchronize(h->stream));
    for (int r = 0; r < h->R; ++r) {
        if (compliance && !(compliance[r] >= 0.0 && compliance[r] <= 1.0)) return fail(h, DYS_ERR_INVALID_ARG, "compliance must be in [0, 1]");
        if (norm_factor) h->rep[(size_t)r].norm_factor = norm_factor[r];
        if (compliance) h->rep[(size_t)r].compliance = compliance[r];
        if (seed) h->rep[(size_t)r].seed = seed[r];
    }
    CK(h, cudaMemcpyAsync(h->d_rep, h->rep.data(), sizeof(RepParam) * (size_t)h->R, cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    if (h->R == 1) { h->c.norm_factor = h->rep[0].norm_factor; h->c.compliance = h->rep[0].compliance; h->c.seed = h->rep[0].seed; }
    return DYS_OK;
}

extern "C" int dys_select_replica(dys_handle* h, int32_t r)
{
    if (!h) return DYS_ERR_INVALID_ARG;
    if (r < -1 || r >= h->R) return fail(h, DYS_ERR_INVALID_ARG, "replica %d out of range (%d replicas)", r, h->R);
    h->sel = r;
    return DYS_OK;
}

static int set_lockdown(dys_handle* h, int32_t scenario, const uint8_t* bld_class, const int32_t* bld_area, const char* who)
{
    if (!h) return DYS_ERR_INVALID_ARG;
    if (!h->have_pop) return fail(h, DYS_ERR_STATE, "%s before dys_load_population", who);
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaStreamSynchronize(h->stream));
    if (scenario == 0) { h->ld.mode = 0; return DYS_OK; }
    if (!bld_class) return fail(h, DYS_ERR_INVALID_ARG, "%s: null bld_class", who);
    if (scenario & ~(SC_GRADUAL | SC_ALL | SC_EDU | SC_REL)) return fail(h, DYS_ERR_INVALID_ARG, "%s: unknown scenario bits", who);
    if (h->p.world > 1) return fail(h, DYS_ERR_NOT_AVAILABLE, "the device lockdown policy needs the GLOBAL R: not in sharded mode (use the host policy)");
    if (h->c.diagnosis < 2 || h->c.diagnosis > 12) return fail(h, DYS_ERR_NOT_AVAILABLE, "the device lockdown policy needs 2 <= diagnosis <= 12");
    if (h->steps != 0 && h->last_day >= 0) return fail(h, DYS_ERR_STATE, "%s after the run has started: reload the state first", who);
    if (h->days_grid < 2) return fail(h, DYS_ERR_NOT_AVAILABLE, "the device lockdown policy needs the cooperative k_days launch");
    DevCtx& c = h->c;
    const int64_t R = h->R;
    int rc = DYS_OK;
    if (bld_area) {
        if (!(h->p.flags & DYS_WANT_SA_HIST)) return fail(h, DYS_ERR_NOT_AVAILABLE, "%s: the per-area policy reads the per-area histogram (DYS_WANT_SA_HIST)", who);
        if (!(scenario & (SC_ALL | SC_EDU | SC_REL))) return fail(h, DYS_ERR_INVALID_ARG, "%s: no building class to close", who);
        for (int64_t b = 0; b < c.B; ++b)
            if (bld_area[b] < 0 || bld_area[b] >= c.S) return fail(h, DYS_ERR_INVALID_ARG, "%s: bld_area[%lld] = %d outside [0, %lld)", who, (long long)b, bld_area[b], (long long)c.S);
    }
    if (!h->ld.flags) {
        uint8_t* cls = nullptr;
        rc = dev_alloc(h, &cls, c.B);
        if (rc == DYS_OK) rc = dev_alloc(h, &h->ld.flags, 4 * R * c.B);
        if (rc == DYS_OK) rc = dev_alloc(h, &h->ld.n_closed, 4 * R);
        if (rc == DYS_OK) rc = dev_alloc(h, &h->ld.r_hist, 16 * R);
        if (rc != DYS_OK) return rc;
        h->ld.bld_class = cls;
    }
    if (bld_area && !h->ld.r_sa) {
        int32_t* area = nullptr;
        rc = dev_alloc(h, &area, c.B);
        if (rc == DYS_OK) rc = dev_alloc(h, &h->ld.r_sa, 16 * R * c.S);
        if (rc == DYS_OK) rc = dev_alloc(h, &h->ld.what_sa, R * c.S);
        if (rc != DYS_OK) return rc;
        h->ld.bld_area = area;
        h->ld.S = (int32_t)c.S;
    }
    rc = copy_in(h, (void*)h->ld.bld_class, bld_class, (size_t)c.B);
    if (rc == DYS_OK && bld_area) rc = copy_in(h, (void*)h->ld.bld_area, bld_area, sizeof(int32_t) * (size_t)c.B);
    if (rc != DYS_OK) return rc;
    h->ld.scen = scenario | (bld_area ? SC_DIFF : 0);
    h->ld.mode = 1;
    return init_lockdown(h);
}

extern "C" int dys_set_lockdown(dys_handle* h, int32_t scenario, const uint8_t* bld_class)
{
    return set_lockdown(h, scenario, bld_class, nullptr, "dys_set_lockdown");
}

extern "C" int dys_set_lockdown_areas(dys_handle* h, int32_t scenario, const uint8_t* bld_class, const int32_t* bld_area)
{
    if (h && scenario != 0 && !bld_area) return fail(h, DYS_ERR_INVALID_ARG, "dys_set_lockdown_areas: null bld_area");
    return set_lockdown(h, scenario, bld_class, bld_area, "dys_set_lockdown_areas");
}

extern "C" int dys_abort(dys_handle* h)
{
    if (!h || !h->h_abort) return DYS_ERR_INVALID_ARG;
    *h->h_abort = 1;
    // the word the kernels poll lives in device memory: a copy on the upload stream reaches it while a launch is running
    static const int32_t one = 1;
    if (h->up_stream && h->c.abort_word) {
        cudaSetDevice(h->device);
        cudaMemcpyAsync((void*)h->c.abort_word, &one, sizeof one, cudaMemcpyHostToDevice, h->up_stream);
    }
    return DYS_OK;
}

extern "C" int dys_load_graph(dys_handle* h, const int64_t* rowptr, const int32_t* col, const double* val)
{
    if (!h) return DYS_ERR_INVALID_ARG;
    if (!h->have_pop) return fail(h, DYS_ERR_STATE, "dys_load_graph before dys_load_population");
    if (!rowptr) return fail(h, DYS_ERR_INVALID_ARG, "dys_load_graph: null rowptr");
    CK(h, cudaSetDevice(h->device));
    DevCtx& c = h->c;
    // staging copies of the caller's CSR (row = susceptible side); freed once the transposed form exists
    int64_t* d_rowptr = nullptr;
    int32_t *d_col = nullptr, *d_cnt = nullptr;
    double* d_val = nullptr;
    unsigned long long* d_cursor = nullptr;
    auto cleanup = [&]() { cudaFree(d_rowptr); cudaFree(d_col); cudaFree(d_val); cudaFree(d_cnt); cudaFree(d_cursor); };
    CK(h, cudaMalloc((void**)&d_rowptr, sizeof(int64_t) * (size_t)(c.n_loc + 1)));
    int rc = copy_in(h, d_rowptr, rowptr, sizeof(int64_t) * (size_t)(c.n_loc + 1));
    int64_t ends[2] = {0, 0};
    if (rc == DYS_OK && (cudaMemcpyAsync(&ends[0], d_rowptr, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
                         cudaMemcpyAsync(&ends[1], d_rowptr + c.n_loc, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
                         cudaStreamSynchronize(h->stream) != cudaSuccess))
        rc = fail(h, DYS_ERR_CUDA, "dys_load_graph: reading rowptr failed");
    if (rc == DYS_OK && (ends[0] != 0 || ends[1] < 0)) rc = fail(h, DYS_ERR_RANGE, "dys_load_graph: rowptr[0] must be 0 and rowptr[N] >= 0");
    const int64_t nnz = ends[1];
    if (rc == DYS_OK && nnz > 0 && (!col || !val)) rc = fail(h, DYS_ERR_INVALID_ARG, "dys_load_graph: null col/val");
    if (rc != DYS_OK) { cleanup(); return rc; }
    const size_t nz = (size_t)(nnz > 0 ? nnz : 1);
    cudaError_t e = cudaMalloc((void**)&d_col, sizeof(int32_t) * nz);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_val, sizeof(double) * nz);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_cnt, sizeof(int32_t) * (size_t)c.rt_n);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_cursor, sizeof(unsigned long long) * (size_t)c.rt_n);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_cnt, 0, sizeof(int32_t) * (size_t)c.rt_n, h->stream);
    if (e != cudaSuccess) { cleanup(); return fail(h, DYS_ERR_NOMEM, "dys_load_graph: %s", cudaGetErrorString(e)); }
    rc = copy_in(h, d_col, col, sizeof(int32_t) * (size_t)nnz);
    if (rc == DYS_OK) rc = copy_in(h, d_val, val, sizeof(double) * (size_t)nnz);
    int64_t* d_tptr = nullptr;
    TEdge* d_tedge = nullptr;
    if (rc == DYS_OK) rc = dev_alloc(h, &d_tptr, c.rt_n + 1);
    if (rc == DYS_OK) {
        k_validate_graph_count<<<grid_for(c.n_loc * 32, 256), 256, 0, h->stream>>>(c, d_rowptr, d_col, d_cnt, h->d_bad);
        rc = check_bad(h, "dys_load_graph");
    }
    if (rc == DYS_OK) {
        // transpose: column counts -> offsets -> scatter with the static shared-building class of each pair
        k_scan_counts<<<1, 1024, 0, h->stream>>>(d_cnt, c.rt_n, d_tptr, d_cursor);
        int64_t kept = 0;      // the non-zeros whose pair can ever share a building
        if (cudaMemcpyAsync(&kept, d_tptr + c.rt_n, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
            cudaStreamSynchronize(h->stream) != cudaSuccess)
            rc = fail(h, DYS_ERR_CUDA, "dys_load_graph: %s", cudaGetErrorString(cudaGetLastError()));
        if (rc == DYS_OK) rc = dev_alloc(h, &d_tedge, kept + 4);
        h->nnz_kept = kept;
    }
    if (rc == DYS_OK) {
        k_fill_transpose<<<grid_for(c.n_loc * 32, 256), 256, 0, h->stream>>>(c, d_rowptr, d_col, d_val, d_cursor, d_tedge);
        if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(h->stream) != cudaSuccess)
            rc = fail(h, DYS_ERR_CUDA, "building the transposed network failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    cleanup();
    if (rc != DYS_OK) { dev_free(h, d_tptr); dev_free(h, d_tedge); return rc; }
    dev_free(h, c.tptr); dev_free(h, c.tedge);      // a previously loaded network
    c.tptr = d_tptr; c.tedge = d_tedge;
    h->nnz = nnz;
    h->have_graph = true;
    return DYS_OK;
}

extern "C" int dys_load_quirk(dys_handle* h, const uint8_t* hh_always, const int64_t* trig_ptr, const int32_t* trig_hh)
{
    if (!h) return DYS_ERR_INVALID_ARG;
    if (!h->have_pop) return fail(h, DYS_ERR_STATE, "dys_load_quirk before dys_load_population");
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaStreamSynchronize(h->stream));
    DevCtx& c = h->c;
    int rc = DYS_OK;
    if (hh_always && h->p.world > 1)
        return fail(h, DYS_ERR_NOT_AVAILABLE, "hh_always (households quarantined whenever ANYONE is diagnosed, contagious3.py:267) "
                    "needs a global diagnosis flag and is not available in sharded mode");
    dev_free(h, c.hh_always); dev_free(h, c.trig_ptr); dev_free(h, c.trig_hh);      // previously loaded tables
    c.hh_always = nullptr; c.trig_ptr = nullptr; c.trig_hh = nullptr;
    if (hh_always) {
        uint8_t* d = nullptr;
        rc = dev_alloc(h, &d, c.H);
        if (rc == DYS_OK) rc = copy_in(h, d, hh_always, (size_t)c.H);
        if (rc != DYS_OK) return rc;
        c.hh_always = d;
    }
    if (trig_ptr) {
        int64_t* d = nullptr;
        rc = dev_alloc(h, &d, c.n_pad + 1);
        if (rc == DYS_OK) rc = copy_in(h, d, trig_ptr, sizeof(int64_t) * (size_t)(c.n_loc + 1));
        if (rc != DYS_OK) return rc;
        int64_t nt = 0;
        CK(h, cudaMemcpyAsync(&nt, d + c.n_loc, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
        CK(h, cudaStreamSynchronize(h->stream));
        if (nt < 0) { dev_free(h, d); return fail(h, DYS_ERR_RANGE, "dys_load_quirk: negative trigger count"); }
        if (nt == 0) { dev_free(h, d); return DYS_OK; }
        if (!trig_hh) { dev_free(h, d); return fail(h, DYS_ERR_INVALID_ARG, "dys_load_quirk: null trig_hh"); }
        if (c.n_pad > c.n_loc) {
            std::vector<int64_t> tail((size_t)(c.n_pad - c.n_loc), nt);
            CK(h, cudaMemcpyAsync(d + c.n_loc + 1, tail.data(), sizeof(int64_t) * tail.size(), cudaMemcpyHostToDevice, h->stream));
            CK(h, cudaStreamSynchronize(h->stream));
        }
        int32_t* t = nullptr;
        rc = dev_alloc(h, &t, nt);
        if (rc == DYS_OK) rc = copy_in(h, t, trig_hh, sizeof(int32_t) * (size_t)nt);
        if (rc == DYS_OK) {
            const int64_t m = c.n_loc > nt ? c.n_loc : nt;
            k_validate_quirk<<<grid_for(m, 256), 256, 0, h->stream>>>(d, t, c.n_loc, nt, c.H, h->d_bad);
            rc = check_bad(h, "dys_load_quirk");
        }
        if (rc != DYS_OK) { dev_free(h, d); dev_free(h, t); return rc; }
        c.trig_ptr = d; c.trig_hh = t;
    }
    CK(h, cudaStreamSynchronize(h->stream));
    return DYS_OK;
}

extern "C" int dys_set_open(dys_handle* h, const uint8_t* open_flags)
{
    if (!h || !open_flags) return h ? fail(h, DYS_ERR_INVALID_ARG, "dys_set_open: null") : DYS_ERR_INVALID_ARG;
    CK(h, cudaSetDevice(h->device));
    DevCtx& c = h->c;
    cudaPointerAttributes at{};
    const bool on_device = cudaPointerGetAttributes(&at, open_flags) == cudaSuccess && at.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    if (on_device) {      // a device-resident flag vector (e.g. a torch CUDA tensor): bring it to the pinned shadow
        CK(h, cudaStreamSynchronize(h->stream));
        CK(h, cudaMemcpy(h->h_open, open_flags, (size_t)c.B, cudaMemcpyDeviceToHost));
    } else {
        // the reference re-assigns build[:,10] every day (contagious3.py:384) but it rarely changes: an unchanged
        // vector costs one memcmp and no transfer
        if (memcmp(h->h_open, open_flags, (size_t)c.B) == 0) return DYS_OK;
        CK(h, cudaStreamSynchronize(h->stream));      // the shadow may still be the source of a queued copy
        memcpy(h->h_open, open_flags, (size_t)c.B);
    }
    int32_t closed = 0;
    for (int64_t b = 0; b < c.B; ++b) closed += h->h_open[b] == 0;
    c.n_closed = closed;
    CK(h, cudaMemcpyAsync((void*)c.open, h->h_open, (size_t)c.B, cudaMemcpyHostToDevice, h->stream));
    return DYS_OK;
}

// the day-independent peer fields of the context: who waits for whom, and which part of the owned slice each write
// target needs (only the bytes inside that rank's halo travel)
static inline int64_t shard_pad(const int64_t lo, const int64_t hi) { return lo + round_up(hi - lo, 1024); }

// global agents of rank `owner`'s slice (padded to its 1024 granule) that lie in rank `reader`'s halo, in whole pieces
static void slice_in_halo(const dys_handle* h, int owner, int reader, int64_t& lo, int64_t& n)
{
    const int64_t o0 = h->peer_own[owner][0], o1 = shard_pad(h->peer_own[owner][0], h->peer_own[owner][1]);
    int64_t a = o0 > h->peer_halo[reader][0] ? o0 : h->peer_halo[reader][0];
    int64_t b = o1 < h->peer_halo[reader][1] ? o1 : h->peer_halo[reader][1];
    if (b <= a) { lo = 0; n = 0; return; }
    a &= ~(int64_t)127;
    b = (b + 127) & ~(int64_t)127;
    if (b > o1) b = o1;
    lo = a; n = b - a;
}

static void bind_peers(dys_handle* h)
{
    DevCtx& c = h->c;
    c.n_peers = h->n_peers; c.rank = h->p.rank;
    c.arrive = h->d_arrive;
    c.pub_lo[0] = 0; c.pub_n[0] = 0; c.units_out[0] = 0; c.arrive_at[0] = h->d_arrive;
    for (int r = 0; r < 8; ++r) c.expect_in[r] = 0;
    int p = 1;
    const int me = h->p.rank;
    for (int r = 0; r < h->n_peers; ++r) {
        for (int k = 0; k < 2; ++k) h->buf.peer_ib[r][k] = (h->n_peers > 1 ? h->peer_ib[r] : h->d_ib) + (size_t)k * h->ib_stride;
        if (h->n_peers > 1 && r != me && ((h->partner_mask >> r) & 1u)) {
            int64_t lo, n;
            slice_in_halo(h, me, r, lo, n);          // what this rank delivers to r
            c.pub_lo[p] = lo; c.pub_n[p] = n;
            c.units_out[p] = (uint32_t)(n >> 7) + 1u;
            c.arrive_at[p] = h->peer_arrive[r] + me;
            slice_in_halo(h, r, me, lo, n);          // what r delivers here
            c.expect_in[r] = (uint32_t)(n >> 7) + 1u;
            ++p;
        }
    }
    c.epoch = h->pub_epoch;
    // boundary split: every pub range must be a prefix or a suffix of the owned range (contiguous shards of communities)
    const int64_t n_pieces = c.n_pad >> 7;
    int64_t a = 0, b = n_pieces;
    bool ok = h->n_peers > 1 && p > 1;
    for (int q = 1; q < p && ok; ++q) {
        const int64_t lo = (c.pub_lo[q] - c.lo) >> 7, hi = lo + (c.pub_n[q] >> 7);
        if (c.pub_n[q] == 0) continue;
        if (lo == 0) a = hi > a ? hi : a;
        else if (hi == n_pieces) b = lo < b ? lo : b;
        else ok = false;
    }
    if (a >= b) ok = false;                        // (the whole slice is boundary: nothing to overlap with)
    c.bnd_ok = ok ? 1 : 0; c.bnd_a = (int32_t)a; c.bnd_b = (int32_t)b;
}

// phase A on its own for `day`: the day's attention bytes / quirk flags / flag block start clean, k_snapshot raises
// into them and writes the snapshot bytes of parity day & 1 (published to the peers in peer mode)
static int run_snapshot(dys_handle* h, int32_t day)
{
    DevCtx& c = h->c;
    const size_t R = (size_t)h->R;
    CK(h, cudaMemsetAsync(h->buf.att[day & 1], 0, ((size_t)c.n_pad + 16) * R, h->stream));
    CK(h, cudaMemsetAsync(h->buf.qflag[day % Q_RING], 0, ((size_t)c.H + 16) * R, h->stream));
    CK(h, cudaMemsetAsync(h->buf.df + (size_t)(day & (DF_RING - 1)) * DF_LEN * R, 0, sizeof(int32_t) * DF_LEN * R, h->stream));
    bind_day(c, h->buf, day); c.bounds_plan = nullptr; c.piece_work = nullptr;   // (the daily cut belongs to k_days)
    bind_ib_targets(c, h->buf, day & 1);
    c.stats = nullptr; c.sa_hist = nullptr; c.bcount = nullptr; c.ev = nullptr;
    for (int r = 0; r < h->R; ++r) {
        DevCtx cr = c;
        shift_replica(cr, r);
        cr.norm_factor = h->rep[(size_t)r].norm_factor; cr.compliance = h->rep[(size_t)r].compliance; cr.seed = h->rep[(size_t)r].seed;
        k_snapshot<<<(unsigned)(c.n_pad / 1024), 256, 0, h->stream>>>(cr, day);
    }
    if (h->n_peers > 1) {
        k_publish<<<PUBLISH_BLOCKS, 256, 0, h->stream>>>(c);
        k_signal<<<1, 32, 0, h->stream>>>(c);
        ++h->pub_epoch;
    }
    CK(h, cudaGetLastError());
    bind_day(c, h->buf, day); c.bounds_plan = nullptr; c.piece_work = nullptr;   // (the daily cut belongs to k_days)
    c.epoch = h->pub_epoch;
    h->ib_day = day;
    return DYS_OK;
}

static void tick(dys_handle* h, int which)
{
    if (h->tev.empty()) return;
    cudaEventRecord(h->tev[(size_t)((h->t_end % dys_handle::T_RING) * 4 + which)], h->stream);
}

static int step_begin(dys_handle* h, int32_t day, const dys_injected* inj)
{
    if (!h->have_pop || !h->have_graph) return fail(h, DYS_ERR_STATE, "dys_step before dys_load_population/dys_load_graph");
    if (h->begun) return fail(h, DYS_ERR_STATE, "dys_step_begin called twice without dys_step_end");
    if (day < 0 || day > 32000) return fail(h, DYS_ERR_INVALID_ARG, "day out of range");
    if (h->R > 1) return fail(h, DYS_ERR_NOT_AVAILABLE, "replicas advance through dys_step_many");
    if (h->ld.mode) return fail(h, DYS_ERR_NOT_AVAILABLE, "the device lockdown policy advances through dys_step_many");
    { int ra = check_abort(h); if (ra != DYS_OK) return ra; }
    CK(h, cudaSetDevice(h->device));
    DevCtx& c = h->c;
    c.u_adm = c.u_death = c.u_pair = nullptr;
    c.pair_prob = nullptr;
    if (inj) {   // injected draws are staged into library-owned device buffers (the caller's may be host memory)
        const double* srcs[3] = {inj->u_adm, inj->u_death, inj->u_pair};
        double** dsts[3] = {&h->d_u_adm, &h->d_u_death, &h->d_u_pair};
        const double** ctxp[3] = {&c.u_adm, &c.u_death, &c.u_pair};
        if (inj->u_pair && c.n_glob > 8192) return fail(h, DYS_ERR_INVALID_ARG, "injected pair draws are limited to N <= 8192");
        for (int k = 0; k < 3; ++k) {
            if (!srcs[k]) continue;
            const int64_t cnt = k == 2 ? c.n_glob * c.n_glob : c.n_glob;
            if (!*dsts[k]) { int rc = dev_alloc(h, dsts[k], cnt, false); if (rc != DYS_OK) return rc; }
            int rc = copy_in(h, *dsts[k], srcs[k], sizeof(double) * (size_t)cnt);
            if (rc != DYS_OK) return rc;
            *ctxp[k] = *dsts[k];
        }
        if (inj->u_pair) {
            if (!h->d_pair_prob) { int rc = dev_alloc(h, &h->d_pair_prob, c.n_glob * c.n_glob); if (rc != DYS_OK) return rc; }
            CK(h, cudaMemsetAsync(h->d_pair_prob, 0, sizeof(double) * (size_t)(c.n_glob * c.n_glob), h->stream));
            c.pair_prob = h->d_pair_prob;
        }
    }
    if (!h->tev.empty() && h->t_end - h->t_begin >= dys_handle::T_RING) h->t_begin = h->t_end - dys_handle::T_RING + 1;
    bind_peers(h);
    bind_day(c, h->buf, day); c.bounds_plan = nullptr; c.piece_work = nullptr;   // (the daily cut belongs to k_days)
    tick(h, 0);
    if (h->ib_day != day || (inj && inj->u_adm)) {
        // first day after a load / set_state, a skipped day, or injected admission draws (tomorrow's diagnoses cannot
        // be predicted from an array the caller has not handed over yet)
        int rc = run_snapshot(h, day);
        if (rc != DYS_OK) return rc;
    }
    tick(h, 1);
    CK(h, cudaGetLastError());
    h->begun = true;
    return DYS_OK;
}

// the slot's previous day keeps its stats in the history; its block (host and device side) is then free
static int retire_slot(dys_handle* h, int slot)
{
    if (h->out_day[slot] >= 0) {
        CK(h, cudaEventSynchronize(h->out_ev[h->out_ev_of[slot]]));
        const int hs = (int)(h->out_step[slot] % DYS_STATS_RING);
        for (int r = 0; r < h->R; ++r)
            memcpy(&h->hist_stats[((size_t)hs * h->R + r) * N_STATS], h->h_out + ((size_t)slot * h->R + r) * h->out_copy_bytes,
                   sizeof(int64_t) * N_STATS);
        h->hist_day[hs] = h->out_day[slot];
        h->out_day[slot] = -1;
    }
    return DYS_OK;
}

// the day's block (one per replica) leaves on the copy stream, under the following days' kernels
static int ship_day(dys_handle* h, int32_t day, int slot, int64_t step)
{
    CK(h, cudaStreamWaitEvent(h->copy_stream, h->day_done[slot], 0));
    unsigned char* dst = h->h_out + (size_t)slot * h->R * h->out_copy_bytes;
    if (h->R == 1) CK(h, cudaMemcpyAsync(dst, h->d_out[slot], h->out_copy_bytes, cudaMemcpyDeviceToHost, h->copy_stream));
    else CK(h, cudaMemcpy2DAsync(dst, h->out_copy_bytes, h->d_out[slot], h->out_dev_bytes, h->out_copy_bytes, (size_t)h->R,
                                 cudaMemcpyDeviceToHost, h->copy_stream));
    CK(h, cudaEventRecord(h->out_ev[slot], h->copy_stream));
    h->out_ev_of[slot] = slot;
    h->out_day[slot] = day;
    h->out_step[slot] = step;
    return DYS_OK;
}

// a window of days ships as ONE unit: one event after the launch, one or two 2-D copies (the ring may wrap), one event after
// the copies -- 5 API calls instead of 4 per day (at 1 M agents the host has ~25 us per simulated day to keep the GPU fed)
static int ship_window(dys_handle* h, int32_t first_day, int32_t n_days)
{
    const int s0 = (int)(h->steps % DYS_OUT_RING), last = (int)((h->steps + n_days - 1) % DYS_OUT_RING);
    CK(h, cudaEventRecord(h->day_done[last], h->stream));
    CK(h, cudaStreamWaitEvent(h->copy_stream, h->day_done[last], 0));
    for (int done = 0; done < n_days;) {
        const int s = (s0 + done) % DYS_OUT_RING;
        const int n = (n_days - done) < (DYS_OUT_RING - s) ? (n_days - done) : (DYS_OUT_RING - s);
        CK(h, cudaMemcpy2DAsync(h->h_out + (size_t)s * h->R * h->out_copy_bytes, h->out_copy_bytes, h->d_out[s], h->out_dev_bytes,
                                h->out_copy_bytes, (size_t)n * h->R, cudaMemcpyDeviceToHost, h->copy_stream));
        done += n;
    }
    CK(h, cudaEventRecord(h->out_ev[last], h->copy_stream));
    for (int k = 0; k < n_days; ++k) {
        const int slot = (s0 + k) % DYS_OUT_RING;
        h->out_ev_of[slot] = last;
        h->out_day[slot] = first_day + k;
        h->out_step[slot] = h->steps + k;
        h->last_slot = slot;
    }
    return DYS_OK;
}

static int step_end(dys_handle* h, int32_t day)
{
    if (!h->begun) return fail(h, DYS_ERR_STATE, "dys_step_end without dys_step_begin");
    CK(h, cudaSetDevice(h->device));
    DevCtx& c = h->c;
    const int slot = (int)(h->steps % DYS_OUT_RING);
    int rc = retire_slot(h, slot);
    if (rc != DYS_OK) return rc;
    c.stats = (int64_t*)h->d_out[slot];
    c.sa_hist = (h->p.flags & DYS_WANT_SA_HIST) ? (int32_t*)(h->d_out[slot] + h->off_sa) : nullptr;
    c.bcount = (h->p.flags & DYS_WANT_BUILDING_COUNTS) ? (int32_t*)(h->d_out[slot] + h->off_bc) : nullptr;
    c.ev = (h->p.flags & DYS_WANT_EVENTS) ? (int2*)(h->d_out[slot] + h->off_ev) : nullptr;
    h->last_slot = slot;
    bind_peers(h);
    bind_day(c, h->buf, day); c.bounds_plan = nullptr; c.piece_work = nullptr;   // (the daily cut belongs to k_days)
    c.epoch = h->pub_epoch;      // (the bytes of `day` are the latest publication)
    k_push<<<(unsigned)h->push_grid, BLOCK, 0, h->stream>>>(c, day);
    tick(h, 2);
    {
        DevCtx ca = c;
        ca.n_wr = 1;             // (stand-alone: the agent pass writes its own buffer, k_publish ships the halo parts)
        k_agent<<<(unsigned)(c.n_pad / 1024), 256, 0, h->stream>>>(ca, day);
    }
    tick(h, 3);
    if (h->n_peers > 1) {        // tomorrow's bytes go to every partner GPU, then the day's units
        k_publish<<<PUBLISH_BLOCKS, 256, 0, h->stream>>>(c);
        k_signal<<<1, 32, 0, h->stream>>>(c);
        ++h->pub_epoch;
    }
    CK(h, cudaGetLastError());
    CK(h, cudaEventRecord(h->day_done[slot], h->stream));
    rc = ship_day(h, day, slot, h->steps);
    if (rc != DYS_OK) return rc;
    h->last_day = day;
    h->ib_day = day + 1;         // k_agent wrote tomorrow's snapshot and household flags
    ++h->steps;
    if (!h->tev.empty()) ++h->t_end;
    h->begun = false;
    return DYS_OK;
}

// n_days in ONE cooperative launch (k_days); the per-day blocks are shipped afterwards
static int step_fused(dys_handle* h, int32_t first_day, int32_t n_days, const uint8_t* open_flags)
{
    CK(h, cudaSetDevice(h->device));
    DevCtx& c = h->c;
    DayPlan plan{};
    plan.first_day = first_day; plan.n_days = n_days;
    plan.off_sa = h->off_sa; plan.off_bc = h->off_bc; plan.off_ev = h->off_ev;
    plan.want_sa = (h->p.flags & DYS_WANT_SA_HIST) != 0;
    plan.want_bc = (h->p.flags & DYS_WANT_BUILDING_COUNTS) != 0;
    plan.want_ev = (h->p.flags & DYS_WANT_EVENTS) != 0;
    plan.test_hang_block = -1;
    if (h->test_hooks) { const char* t = getenv("DYS_TEST_HANG_BLOCK"); plan.test_hang_block = t ? atoi(t) : -1; }
    bind_peers(h);
    plan.buf = h->buf;
    // Open flags: the caller's vectors are this window's host input.  They are staged in one of two pinned rings (by
    // window parity) and uploaded on their own stream, so the flags of window w + 1 travel while window w runs; the
    // kernels of this window wait for the upload, the upload waits for the launch that last read the ring.
    const uint8_t* cur_dev = c.open;
    int32_t cur_closed = c.n_closed;
    const int wp = (int)(h->windows & 1);
    bool ring_touched = false;
    plan.ld = h->ld;
    if (h->ld.mode) {
        if (open_flags) return fail(h, DYS_ERR_INVALID_ARG, "the device lockdown policy decides the flags: open_flags must be NULL");
        if (h->last_day < 0 && first_day != 0) return fail(h, DYS_ERR_STATE, "the device lockdown policy starts at day 0");
        if (h->last_day >= 0 && first_day != h->last_day + 1) return fail(h, DYS_ERR_STATE, "the device lockdown policy needs consecutive days");
    }
    if (open_flags) {
        // only the days whose vector differs from the one in force before them travel: a lockdown scenario changes its
        // flags on a handful of days, and a metro's vector is half a megabyte (staging 7 of them per window made the
        // HOST the bottleneck of a 10 M-agent run)
        const size_t B = (size_t)c.B;
        if (h->ring_used[wp]) CK(h, cudaEventSynchronize(h->up_done[wp]));      // (two windows ago: long done)
        bool waited = !h->ring_used[wp];
        int uploads = 0;
        for (int k = 0; k < n_days; ++k) {
            const uint8_t* f = open_flags + (size_t)k * B;
            if (memcmp(h->h_open, f, B) != 0 || h->upload_always) {
                uint8_t* stage = h->h_open_ring[wp] + (size_t)k * B;
                memcpy(stage, f, B);
                memcpy(h->h_open, f, B);
                int32_t closed = 0;
                for (size_t b = 0; b < B; ++b) closed += f[b] == 0;
                if (!waited) { CK(h, cudaStreamWaitEvent(h->up_stream, h->ring_free[wp], 0)); waited = true; }
                CK(h, cudaMemcpyAsync(h->d_open_ring[wp] + (size_t)k * B, stage, B, cudaMemcpyHostToDevice, h->up_stream));
                h->h2d_bytes += (int64_t)B;
                cur_dev = h->d_open_ring[wp] + (size_t)k * B;
                cur_closed = closed;
                ++uploads;
            }
            plan.open[k] = cur_dev;
            plan.n_closed[k] = cur_closed;
        }
        if (uploads) {
            CK(h, cudaEventRecord(h->up_done[wp], h->up_stream));
            CK(h, cudaStreamWaitEvent(h->stream, h->up_done[wp], 0));
        }
        ring_touched = uploads > 0;
    } else {
        for (int k = 0; k < n_days; ++k) { plan.open[k] = cur_dev; plan.n_closed[k] = cur_closed; }
    }
    int rc = DYS_OK;
    for (int k = 0; k < n_days && rc == DYS_OK; ++k) {
        const int slot = (int)((h->steps + k) % DYS_OUT_RING);
        rc = retire_slot(h, slot);
        plan.out[k] = h->d_out[slot];
    }
    if (rc != DYS_OK) return rc;
    c.u_adm = c.u_death = c.u_pair = nullptr; c.pair_prob = nullptr;
    if (h->ib_day != first_day) {      // first day's snapshot, if the previous day did not leave it
        rc = run_snapshot(h, first_day);
        if (rc != DYS_OK) return rc;
    }
    bind_day(c, h->buf, first_day);    // (k_days binds every day's view itself)
    CK(h, cudaMemsetAsync(h->d_barrier, 0, sizeof(unsigned long long) * (size_t)(2 + h->days_grid), h->stream));
    plan.work = h->no_steal ? nullptr : h->d_work;
    // (the service block plans the cut alone: beyond a few staging chunks of pieces it would be late for the next barrier,
    // and the equal split + stealing balances as well -- profiles/r02e_ab.txt)
    plan.plan_max_pieces = h->plan_max;
    plan.epoch0 = h->pub_epoch;
    if (h->n_peers > 1) h->pub_epoch += (unsigned long long)n_days;      // every day of the window publishes tomorrow's bytes
    unsigned int* bar = h->d_barrier;
    void* args[] = {(void*)&c, (void*)&plan, (void*)&bar};
    const size_t trace_words = (size_t)MAX_FUSED_DAYS * (size_t)h->days_grid * TRACE_SLOTS;
    if (h->trace_path && h->trace_path[0]) {
        if (!h->d_trace) { rc = dev_alloc(h, &h->d_trace, (int64_t)trace_words); if (rc != DYS_OK) return rc; }
        plan.trace = h->d_trace;
    }
    // the kernel variant: PEERS carries the exchange code, MULTI the replica views and the work stealing of long runs
    const bool multi = h->R > 1 || (!h->no_steal && (c.n_pad >> 7) > 32ll * (h->days_grid - 1));
    const void* kern = h->n_peers > 1 ? (multi ? (const void*)k_days<true, true> : (const void*)k_days<true, false>)
                                      : (multi ? (const void*)k_days<false, true> : (const void*)k_days<false, false>);
    CK(h, cudaLaunchCooperativeKernel(kern, dim3((unsigned)h->days_grid), dim3(256), args, 0, h->stream));

    if (plan.trace) {      // developer tracing: synchronous, appended as text
        std::vector<unsigned long long> tr(trace_words);
        CK(h, cudaMemcpyAsync(tr.data(), h->d_trace, trace_words * 8, cudaMemcpyDeviceToHost, h->stream));
        CK(h, cudaStreamSynchronize(h->stream));
        if (FILE* f = fopen(h->trace_path, "a")) {
            for (int k = 0; k < n_days; ++k)
                for (int b = 0; b < h->days_grid; ++b) {
                    const unsigned long long* t = &tr[((size_t)k * h->days_grid + b) * TRACE_SLOTS];
                    fprintf(f, "%d %d", first_day + k, b);
                    for (int i = 0; i < TRACE_SLOTS; ++i) fprintf(f, " %llu", t[i]);
                    fprintf(f, "\n");
                }
            fclose(f);
        }
    }
    if (cur_dev != c.open) {      // the flags in force after the window become the standing vector
        CK(h, cudaMemcpyAsync((void*)c.open, cur_dev, (size_t)c.B, cudaMemcpyDeviceToDevice, h->stream));
        c.n_closed = cur_closed;
    }
    if (ring_touched) {           // (after the copy above: it still reads the ring)
        CK(h, cudaEventRecord(h->ring_free[wp], h->stream));
        h->ring_used[wp] = true;
        ++h->windows;
    }
    rc = ship_window(h, first_day, n_days);
    if (rc != DYS_OK) return rc;
    h->steps += n_days;
    h->last_day = first_day + n_days - 1;
    h->ib_day = first_day + n_days;
    return DYS_OK;
}

extern "C" int dys_step_begin(dys_handle* h, int32_t day, const dys_injected* injected)
{
    return h ? step_begin(h, day, injected) : DYS_ERR_INVALID_ARG;
}

extern "C" int dys_step_end(dys_handle* h, int32_t day) { return h ? step_end(h, day) : DYS_ERR_INVALID_ARG; }

extern "C" int dys_step(dys_handle* h, int32_t day, const dys_injected* injected)
{
    if (!h) return DYS_ERR_INVALID_ARG;
    int rc = step_begin(h, day, injected);
    return rc == DYS_OK ? step_end(h, day) : rc;
}

static int find_out_slot(dys_handle* h, int32_t day)
{
    for (int i = 0; i < DYS_OUT_RING; ++i)
        if (h->out_day[i] == day) return i;
    return -1;
}

extern "C" int dys_step_many(dys_handle* h, int32_t first_day, int32_t n_days, const uint8_t* open_flags)
{
    if (!h) return DYS_ERR_INVALID_ARG;
    if (n_days < 1 || n_days > DYS_OUT_RING - 2) return fail(h, DYS_ERR_INVALID_ARG, "dys_step_many: 1..%d days", DYS_OUT_RING - 2);
    if (!h->have_pop || !h->have_graph) return fail(h, DYS_ERR_STATE, "dys_step_many before dys_load_population/dys_load_graph");
    if (h->begun) return fail(h, DYS_ERR_STATE, "dys_step_many between dys_step_begin and dys_step_end");
    if (first_day < 0 || first_day + n_days > 32000) return fail(h, DYS_ERR_INVALID_ARG, "day out of range");
    { int ra = check_abort(h); if (ra != DYS_OK) return ra; }
    if ((n_days >= 2 || h->R > 1 || h->ld.mode) && n_days <= MAX_FUSED_DAYS && h->days_grid > 0 && h->tev.empty() &&
        (!h->no_fuse || h->R > 1 || h->ld.mode))
        return step_fused(h, first_day, n_days, open_flags);
    if (h->ld.mode) return fail(h, DYS_ERR_NOT_AVAILABLE, "the device lockdown policy runs inside dys_step_many windows of <= %d days", MAX_FUSED_DAYS);
    for (int32_t k = 0; k < n_days; ++k) {
        if (open_flags) {
            int rc = dys_set_open(h, open_flags + (size_t)k * (size_t)h->c.B);
            if (rc != DYS_OK) return rc;
        }
        int rc = step_begin(h, first_day + k, nullptr);
        if (rc == DYS_OK) rc = step_end(h, first_day + k);
        if (rc != DYS_OK) return rc;
    }
    return DYS_OK;
}

static inline int sel_rep(const dys_handle* h) { return h->sel >= 0 ? h->sel : 0; }

extern "C" int dys_get_outputs(dys_handle* h, int32_t day, dys_day_outputs* out)
{
    if (!h || !out) return DYS_ERR_INVALID_ARG;
    const int i = find_out_slot(h, day);
    if (i < 0) return fail(h, DYS_ERR_NOT_AVAILABLE, "outputs of day %d are not in the ring (the last %d days are kept)", day, DYS_OUT_RING);
    CK(h, cudaEventSynchronize(h->out_ev[h->out_ev_of[i]]));
    { int ra = check_abort(h); if (ra != DYS_OK) return ra; }
    const unsigned char* base = h->h_out + ((size_t)i * h->R + sel_rep(h)) * h->out_copy_bytes;
    out->day = day;
    out->stats = (const int64_t*)base;
    out->sa_hist = (h->p.flags & DYS_WANT_SA_HIST) ? (const int32_t*)(base + h->off_sa) : nullptr;
    out->building_counts = (h->p.flags & DYS_WANT_BUILDING_COUNTS) ? (const int32_t*)(base + h->off_bc) : nullptr;
    out->events = (h->p.flags & DYS_WANT_EVENTS) ? (const int32_t*)(base + h->off_ev) : nullptr;
    out->n_events = out->stats[ST_N_EVENTS];
    out->events_inline = (int32_t)(out->n_events < h->ev_inline ? out->n_events : h->ev_inline);
    out->events_capacity = h->ev_inline;
    return DYS_OK;
}

extern "C" int dys_get_stats(dys_handle* h, int32_t day, int64_t* out)
{
    if (!h || !out) return DYS_ERR_INVALID_ARG;
    const int i = find_out_slot(h, day);
    if (i >= 0) {
        CK(h, cudaEventSynchronize(h->out_ev[h->out_ev_of[i]]));
        { int ra = check_abort(h); if (ra != DYS_OK) return ra; }
        memcpy(out, h->h_out + ((size_t)i * h->R + sel_rep(h)) * h->out_copy_bytes, sizeof(int64_t) * N_STATS);
        return DYS_OK;
    }
    for (int k = 0; k < DYS_STATS_RING; ++k)
        if (h->hist_day[k] == day) {
            memcpy(out, &h->hist_stats[((size_t)k * h->R + sel_rep(h)) * N_STATS], sizeof(int64_t) * N_STATS);
            return DYS_OK;
        }
    return fail(h, DYS_ERR_NOT_AVAILABLE, "stats of day %d are no longer kept (the last %d days are)", day, DYS_STATS_RING);
}

static int copy_out(dys_handle* h, void* dst, const void* src, size_t bytes)
{
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return DYS_OK;
}

extern "C" int dys_get_sa_hist(dys_handle* h, int32_t day, int32_t* out)
{
    if (!h || !out) return DYS_ERR_INVALID_ARG;
    dys_day_outputs o;
    int rc = dys_get_outputs(h, day, &o);
    if (rc != DYS_OK) return rc;
    if (!o.sa_hist) return fail(h, DYS_ERR_NOT_AVAILABLE, "sa_hist was not requested in dys_params.flags");
    memcpy(out, o.sa_hist, sizeof(int32_t) * (size_t)(h->c.S_out * HIST_LEN));
    return DYS_OK;
}

extern "C" int dys_get_building_counts(dys_handle* h, int32_t day, int32_t* out)
{
    if (!h || !out) return DYS_ERR_INVALID_ARG;
    dys_day_outputs o;
    int rc = dys_get_outputs(h, day, &o);
    if (rc != DYS_OK) return rc;
    if (!o.building_counts) return fail(h, DYS_ERR_NOT_AVAILABLE, "building counts were not requested in dys_params.flags");
    memcpy(out, o.building_counts, sizeof(int32_t) * (size_t)h->c.B_out);
    return DYS_OK;
}

extern "C" int dys_get_io_mat(dys_handle* h, int32_t day, int32_t* out)
{
    if (!h || !out) return DYS_ERR_INVALID_ARG;
    if (!h->c.io_mat) return fail(h, DYS_ERR_NOT_AVAILABLE, "io_mat was not requested in dys_params.flags");
    if (day != h->last_day) return fail(h, DYS_ERR_NOT_AVAILABLE, "io_mat is only kept for the most recent day (%d)", h->last_day);
    return copy_out(h, out, h->c.io_mat, sizeof(int32_t) * (size_t)(h->c.S * h->c.S));
}

extern "C" int dys_get_events(dys_handle* h, int32_t day, int32_t* agent, int32_t* infector, int64_t cap, int64_t* n)
{
    if (!h || !n) return DYS_ERR_INVALID_ARG;
    dys_day_outputs o;
    int rc = dys_get_outputs(h, day, &o);
    if (rc != DYS_OK) return rc;
    if (!o.events) return fail(h, DYS_ERR_NOT_AVAILABLE, "events were not requested in dys_params.flags");
    const int64_t cnt = o.n_events;
    *n = cnt;
    if (cnt == 0 || (!agent && !infector)) return DYS_OK;
    if (cap < cnt) return fail(h, DYS_ERR_INVALID_ARG, "dys_get_events: capacity %lld < %lld events", (long long)cap, (long long)cnt);
    const int32_t* pairs = o.events;
    std::vector<int32_t> rest;
    if (cnt > o.events_inline) {
        // more events than travel with the day's block: fetch them from the device block, which is only intact until
        // the day after next is stepped
        const int i = find_out_slot(h, day);
        if (h->steps - h->out_step[i] >= DYS_OUT_RING) return fail(h, DYS_ERR_NOT_AVAILABLE, "events of day %d beyond the first %d were overwritten", day, o.events_inline);
        rest.resize((size_t)cnt * 2);
        // (on the upload stream: the main and copy streams may hold windows queued AFTER this day, which this fetch must
        // not wait for; the day itself is complete -- dys_get_outputs above waited for its block)
        CK(h, cudaMemcpyAsync(rest.data(), h->d_out[i] + (size_t)sel_rep(h) * h->out_dev_bytes + h->off_ev, sizeof(int32_t) * 2 * (size_t)cnt,
                              cudaMemcpyDeviceToHost, h->up_stream));
        CK(h, cudaStreamSynchronize(h->up_stream));
        pairs = rest.data();
    }
    for (int64_t k = 0; k < cnt; ++k) {
        if (agent) agent[k] = pairs[2 * k];
        if (infector) infector[k] = pairs[2 * k + 1];
    }
    return DYS_OK;
}

extern "C" int dys_get_state(dys_handle* h, uint8_t* status_x2, int32_t* t_inf, int32_t* t_q, int32_t* t_adm, int32_t* src)
{
    if (!h) return DYS_ERR_INVALID_ARG;
    if (!h->have_pop) return fail(h, DYS_ERR_STATE, "dys_get_state before dys_load_population");
    CK(h, cudaSetDevice(h->device));
    DevCtx c = h->c;
    { const size_t o = (size_t)c.n_pad * sel_rep(h); c.status += o; c.rec += o; c.src += o; }
    const int64_t n = c.n_loc;
    if (status_x2) CK(h, cudaMemcpyAsync(status_x2, c.status, (size_t)n, cudaMemcpyDefault, h->stream));
    if (src) CK(h, cudaMemcpyAsync(src, c.src, sizeof(int32_t) * (size_t)n, cudaMemcpyDefault, h->stream));
    int32_t* outs[3] = {t_inf, t_q, t_adm};
    int32_t* tmp = nullptr;
    for (int k = 0; k < 3; ++k) {
        if (!outs[k]) continue;
        if (!tmp) CK(h, cudaMalloc((void**)&tmp, sizeof(int32_t) * (size_t)n));
        k_widen_days<<<grid_for(n, 256), 256, 0, h->stream>>>(c.rec, k, tmp, n);
        cudaError_t e = cudaMemcpyAsync(outs[k], tmp, sizeof(int32_t) * (size_t)n, cudaMemcpyDefault, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) { cudaFree(tmp); return fail(h, DYS_ERR_CUDA, "dys_get_state: %s", cudaGetErrorString(e)); }
    }
    CK(h, cudaStreamSynchronize(h->stream));
    if (tmp) cudaFree(tmp);
    return DYS_OK;
}

extern "C" int dys_get_pair_prob(dys_handle* h, double* out)
{
    if (!h || !out) return DYS_ERR_INVALID_ARG;
    if (!h->d_pair_prob) return fail(h, DYS_ERR_NOT_AVAILABLE, "pair probabilities are only recorded in injected mode");
    return copy_out(h, out, h->d_pair_prob, sizeof(double) * (size_t)(h->c.n_glob * h->c.n_glob));
}

// Everything the host loop needs from one day, in ONE call: waits for the day's block, copies the stats and the per-building
// counts out of the pinned ring, and forms R (global and per owned area) from the integer histograms -- compute_R,
// contagious3.py:29-42, in the reference's order of float operations.  (The host loop has ~20 us per simulated day at 1 M
// agents; a Python-level consumer spends that on array bookkeeping alone.)
extern "C" int dys_consume_day(dys_handle* h, int32_t day, const double* pdf, int32_t recover, int64_t* stats /* [64] */,
                               double* R_global /* [1] */, double* R_areas /* [S_out] or NULL */,
                               int32_t* building_counts /* [B_out] or NULL */, int32_t* events /* [cap][2] or NULL */,
                               int64_t events_cap, int64_t* n_events)
{
    if (!h || !pdf || !stats || !R_global) return DYS_ERR_INVALID_ARG;
    dys_day_outputs o;
    int rc = dys_get_outputs(h, day, &o);
    if (rc != DYS_OK) return rc;
    memcpy(stats, o.stats, sizeof(int64_t) * N_STATS);
    if (n_events) *n_events = o.n_events;
    rc = dys_compute_r_rows(o.stats + HIST0, 8, 1, HIST_LEN, pdf, recover, R_global);
    if (rc == DYS_OK && R_areas) {
        if (!o.sa_hist) return fail(h, DYS_ERR_NOT_AVAILABLE, "sa_hist was not requested in dys_params.flags");
        rc = dys_compute_r_rows(o.sa_hist, 4, h->c.S_out, HIST_LEN, pdf, recover, R_areas);
    }
    if (rc == DYS_OK && building_counts) {
        if (!o.building_counts) return fail(h, DYS_ERR_NOT_AVAILABLE, "building counts were not requested in dys_params.flags");
        memcpy(building_counts, o.building_counts, sizeof(int32_t) * (size_t)h->c.B_out);
    }
    if (rc == DYS_OK && events && o.n_events > 0) {
        if (!o.events) return fail(h, DYS_ERR_NOT_AVAILABLE, "events were not requested in dys_params.flags");
        if (o.n_events > events_cap) return fail(h, DYS_ERR_INVALID_ARG, "dys_consume_day: %lld events, room for %lld", (long long)o.n_events, (long long)events_cap);
        if (o.n_events <= o.events_inline) memcpy(events, o.events, sizeof(int32_t) * 2 * (size_t)o.n_events);
        else {      // (more than travel with the block: fetch, de-interleave back into pairs)
            std::vector<int32_t> a((size_t)o.n_events), b((size_t)o.n_events);
            int64_t n = 0;
            rc = dys_get_events(h, day, a.data(), b.data(), o.n_events, &n);
            for (int64_t k = 0; rc == DYS_OK && k < n; ++k) { events[2 * k] = a[(size_t)k]; events[2 * k + 1] = b[(size_t)k]; }
        }
    }
    return rc;
}

extern "C" int dys_sync(dys_handle* h)
{
    if (!h) return DYS_ERR_INVALID_ARG;
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaStreamSynchronize(h->stream));
    CK(h, cudaStreamSynchronize(h->copy_stream));
    return check_abort(h);
}

extern "C" int64_t dys_h2d_bytes(const dys_handle* h) { return h ? h->h2d_bytes : 0; }

// compute_R (contagious3.py:29-42) for `rows` days-since-infection histograms in HOST memory: the float epilogue of the
// integer outputs, in the reference's order of operations (sum_I += n_i * pdf(i) for i = 1 .. recover-1, ascending; IEEE
// double, no contraction), so the values compare with == against the reference's.
extern "C" int dys_compute_r_rows(const void* hist, int32_t elem_bytes, int64_t rows, int64_t row_stride, const double* pdf,
                                  int32_t recover, double* out)
{
    if (!hist || !pdf || !out || rows < 0 || (elem_bytes != 4 && elem_bytes != 8) || recover < 1 || recover > HIST_LEN) return DYS_ERR_INVALID_ARG;
    // (host code of this file is compiled with -ffp-contract=off: every product and every sum is rounded on its own)
    for (int64_t r = 0; r < rows; ++r) {
        double sum_I = 0.0, first;
        if (elem_bytes == 4) {
            const int32_t* hrow = (const int32_t*)hist + r * row_stride;
            first = (double)hrow[0];
            for (int i = 1; i < recover; ++i) sum_I = sum_I + (double)hrow[i] * pdf[i];
        } else {
            const int64_t* hrow = (const int64_t*)hist + r * row_stride;
            first = (double)hrow[0];
            for (int i = 1; i < recover; ++i) sum_I = sum_I + (double)hrow[i] * pdf[i];
        }
        out[r] = sum_I > 0.0 ? first / sum_I : 0.0;
    }
    return DYS_OK;
}

extern "C" int dys_get_kernel_times(dys_handle* h, double* ms_total, int64_t* launches)
{
    if (!h || !ms_total || !launches) return DYS_ERR_INVALID_ARG;
    if (h->tev.empty()) return fail(h, DYS_ERR_NOT_AVAILABLE, "DYS_TIME_KERNELS was not set");
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaStreamSynchronize(h->stream));
    for (int k = 0; k < DYS_N_KERNELS; ++k) { ms_total[k] = 0.0; launches[k] = 0; }
    for (int64_t s = h->t_begin; s < h->t_end; ++s) {
        const size_t b = (size_t)((s % dys_handle::T_RING) * 4);
        for (int k = 0; k < DYS_N_KERNELS; ++k) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, h->tev[b + k], h->tev[b + k + 1]) == cudaSuccess) { ms_total[k] += ms; ++launches[k]; }
        }
    }
    h->t_begin = h->t_end;
    return DYS_OK;
}

extern "C" int dys_device_buffer(dys_handle* h, int which, void** ptr, int64_t* bytes)
{
    if (!h || !ptr || !bytes) return DYS_ERR_INVALID_ARG;
    const DevCtx& c = h->c;
    switch (which) {
        case DYS_BUF_STATUS: *ptr = c.status; *bytes = c.n_pad; break;
        case DYS_BUF_INFECTIOUS: *ptr = h->d_ib; *bytes = 2 * h->ib_stride; break;   /* [2][bytes/2]: buffer of day d is d & 1 */
        case DYS_BUF_STATS: *ptr = h->d_out[h->last_slot]; *bytes = (int64_t)sizeof(int64_t) * N_STATS; break;
        case DYS_BUF_AGENT_REC: *ptr = c.rec; *bytes = c.n_pad * (int64_t)sizeof(AgentRec); break;
        default: return fail(h, DYS_ERR_INVALID_ARG, "unknown buffer %d", which);
    }
    return DYS_OK;
}

extern "C" int dys_cuda_stream(dys_handle* h, void** stream)
{
    if (!h || !stream) return DYS_ERR_INVALID_ARG;
    *stream = (void*)h->stream;
    return DYS_OK;
}

extern "C" int64_t dys_device_bytes(const dys_handle* h) { return h ? h->dev_bytes : 0; }

extern "C" int dys_keyed_uniform_device(int device, uint64_t seed, uint32_t day, uint32_t stream, const uint32_t* a,
                                        const uint32_t* b, int64_t n, double* out)
{
    if (!a || !b || !out || n <= 0) return fail(nullptr, DYS_ERR_INVALID_ARG, "dys_keyed_uniform_device: bad argument");
    uint32_t *da = nullptr, *db = nullptr;
    double* dout = nullptr;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMalloc((void**)&da, 4 * (size_t)n);
    if (e == cudaSuccess) e = cudaMalloc((void**)&db, 4 * (size_t)n);
    if (e == cudaSuccess) e = cudaMalloc((void**)&dout, 8 * (size_t)n);
    if (e == cudaSuccess) e = cudaMemcpy(da, a, 4 * (size_t)n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(db, b, 4 * (size_t)n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        k_keyed_uniform<<<grid_for(n, 256), 256>>>(seed, day, stream, da, db, n, dout);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, dout, 8 * (size_t)n, cudaMemcpyDeviceToHost);
    cudaFree(da); cudaFree(db); cudaFree(dout);
    if (e != cudaSuccess) return fail(nullptr, DYS_ERR_CUDA, "dys_keyed_uniform_device: %s", cudaGetErrorString(e));
    return DYS_OK;
}

// ---- peer mode: the per-day exchange fused into k_agent (P2P stores over NVLink) ----
extern "C" int dys_peer_export(dys_handle* h, void* out /* DYS_PEER_HANDLE_BYTES */)
{
    if (!h || !out) return DYS_ERR_INVALID_ARG;
    CK(h, cudaSetDevice(h->device));
    cudaIpcMemHandle_t hs[2];
    CK(h, cudaIpcGetMemHandle(&hs[0], h->d_ib));
    CK(h, cudaIpcGetMemHandle(&hs[1], h->d_arrive));
    static_assert(sizeof(hs) == DYS_PEER_HANDLE_BYTES, "handle size");
    memcpy(out, hs, sizeof hs);
    return DYS_OK;
}

extern "C" int dys_peer_detach(dys_handle* h);

extern "C" int dys_peer_attach(dys_handle* h, int n_ranks, const void* handles /* [n_ranks][DYS_PEER_HANDLE_BYTES] */,
                               const int64_t* halos /* [n_ranks][4] */)
{
    if (!h || !handles || !halos) return DYS_ERR_INVALID_ARG;
    if (n_ranks < 2 || n_ranks > 8 || n_ranks != h->p.world) return fail(h, DYS_ERR_INVALID_ARG, "dys_peer_attach: 2..8 ranks, equal to dys_params.world");
    CK(h, cudaSetDevice(h->device));
    const cudaIpcMemHandle_t* hs = (const cudaIpcMemHandle_t*)handles;
    for (int r = 0; r < n_ranks; ++r) {
        if (r == h->p.rank) { h->peer_ib[r] = h->d_ib; h->peer_arrive[r] = h->d_arrive; continue; }
        void *pib = nullptr, *pfl = nullptr;
        cudaError_t e1 = cudaIpcOpenMemHandle(&pib, hs[2 * r], cudaIpcMemLazyEnablePeerAccess);
        cudaError_t e2 = e1 == cudaSuccess ? cudaIpcOpenMemHandle(&pfl, hs[2 * r + 1], cudaIpcMemLazyEnablePeerAccess) : e1;
        if (e2 != cudaSuccess) {
            cudaGetLastError();
            h->n_peers = r;               // so that dys_peer_detach closes what was opened
            h->peer_ib[r] = (uint8_t*)pib;
            dys_peer_detach(h);
            return fail(h, DYS_ERR_NOT_AVAILABLE, "peer memory of rank %d is not reachable (%s): use the NCCL exchange", r, cudaGetErrorString(e2));
        }
        h->peer_ib[r] = (uint8_t*)pib;
        h->peer_arrive[r] = (unsigned long long*)pfl;
    }
    h->partner_mask = 0;
    for (int r = 0; r < n_ranks; ++r) {
        h->peer_own[r][0] = halos[4 * r]; h->peer_own[r][1] = halos[4 * r + 1];
        h->peer_halo[r][0] = halos[4 * r + 2]; h->peer_halo[r][1] = halos[4 * r + 3];
    }
    const int me = h->p.rank;
    for (int r = 0; r < n_ranks; ++r) {
        if (r == me) continue;
        // Partners = ranks whose agents lie in my halo OR whose halo holds mine.  The relation is made SYMMETRIC (both
        // sides evaluate the same two tests): a rank that only delivers would otherwise never wait, run days ahead and
        // overwrite the parity buffer its partner is still scanning; with the daily token both sides stay within a day.
        const bool r_in_mine = h->peer_own[r][0] < h->peer_halo[me][1] && h->peer_own[r][1] > h->peer_halo[me][0];
        const bool mine_in_r = h->peer_own[me][0] < h->peer_halo[r][1] && h->peer_own[me][1] > h->peer_halo[r][0];
        if (r_in_mine || mine_in_r) h->partner_mask |= 1u << r;
    }
    h->n_peers = n_ranks;
    h->ib_day = -1;
    bind_peers(h);
    return init_cut(h);          // (the number of work blocks changed)
}

extern "C" int dys_peer_detach(dys_handle* h)
{
    if (!h) return DYS_ERR_INVALID_ARG;
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaStreamSynchronize(h->stream));
    for (int r = 0; r < h->n_peers; ++r) {
        if (r == h->p.rank) continue;
        if (h->peer_ib[r]) cudaIpcCloseMemHandle(h->peer_ib[r]);
        if (h->peer_arrive[r]) cudaIpcCloseMemHandle(h->peer_arrive[r]);
        h->peer_ib[r] = nullptr; h->peer_arrive[r] = nullptr;
    }
    cudaGetLastError();
    h->n_peers = 1;
    h->partner_mask = 0;
    h->ib_day = -1;
    bind_peers(h);
    return init_cut(h);
}

// ---- routine generation (contagious3.py:139-170): no handle needed, host arrays in, host array out ----
extern "C" int dys_generate_routines(int device, int64_t n_agents, int64_t n_bld, int32_t V, int32_t k, const int32_t* home,
                                     const int32_t* anchors, const int64_t* out_ptr, const int32_t* out_idx,
                                     const int64_t* in_ptr, const int32_t* in_idx, const int64_t* near_ptr,
                                     const int32_t* near_idx, uint64_t seed, int32_t* routine)
{
    if (!home || !anchors || !out_ptr || !in_ptr || !near_ptr || !routine || n_agents <= 0 || n_bld <= 0)
        return fail(nullptr, DYS_ERR_INVALID_ARG, "dys_generate_routines: null or empty argument");
    if (k < 0 || V < 1 + 3 * (k + 1) || V > 64) return fail(nullptr, DYS_ERR_INVALID_ARG, "dys_generate_routines: need 1 + 3 (k + 1) <= V <= 64");
    for (int64_t a = 0; a < n_agents; ++a) {
        if (home[a] < 0 || home[a] >= n_bld) return fail(nullptr, DYS_ERR_RANGE, "dys_generate_routines: home building out of range");
        for (int s = 0; s < 3; ++s)
            if (anchors[a * 3 + s] < -1 || anchors[a * 3 + s] >= n_bld) return fail(nullptr, DYS_ERR_RANGE, "dys_generate_routines: anchor out of range");
    }
    const int64_t n_out = out_ptr[n_bld], n_in = in_ptr[n_bld], n_near = near_ptr[n_agents];
    if (n_out < 0 || n_in < 0 || n_near < 0 || (n_out && !out_idx) || (n_in && !in_idx) || (n_near && !near_idx))
        return fail(nullptr, DYS_ERR_INVALID_ARG, "dys_generate_routines: bad list arrays");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(nullptr, DYS_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(nullptr, DYS_ERR_INVALID_ARG, "device %d out of range", device);
    cudaError_t e = cudaSetDevice(device);
    std::vector<void*> tmp;
    auto up = [&](const void* src, size_t bytes) -> void* {
        void* d = nullptr;
        if (e == cudaSuccess) e = cudaMalloc(&d, bytes ? bytes : 1);
        if (e == cudaSuccess) { tmp.push_back(d); if (bytes) e = cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice); }
        return d;
    };
    const int32_t* d_home = (const int32_t*)up(home, 4 * (size_t)n_agents);
    const int32_t* d_anch = (const int32_t*)up(anchors, 12 * (size_t)n_agents);
    const int64_t* d_op = (const int64_t*)up(out_ptr, 8 * (size_t)(n_bld + 1));
    const int32_t* d_oi = (const int32_t*)up(out_idx, 4 * (size_t)n_out);
    const int64_t* d_ip = (const int64_t*)up(in_ptr, 8 * (size_t)(n_bld + 1));
    const int32_t* d_ii = (const int32_t*)up(in_idx, 4 * (size_t)n_in);
    const int64_t* d_np = (const int64_t*)up(near_ptr, 8 * (size_t)(n_agents + 1));
    const int32_t* d_ni = (const int32_t*)up(near_idx, 4 * (size_t)n_near);
    int32_t* d_r = nullptr;
    int32_t* d_bad = nullptr;
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_r, 4 * (size_t)n_agents * V);
    if (e == cudaSuccess) { tmp.push_back(d_r); e = cudaMalloc((void**)&d_bad, 4); }
    if (e == cudaSuccess) { tmp.push_back(d_bad); e = cudaMemset(d_bad, 0, 4); }
    int32_t bad = 0;
    if (e == cudaSuccess) {
        k_validate_lists<<<grid_for(n_bld, 256), 256>>>(d_op, d_oi, n_bld, n_bld, d_bad);
        k_validate_lists<<<grid_for(n_bld, 256), 256>>>(d_ip, d_ii, n_bld, n_bld, d_bad);
        k_validate_lists<<<grid_for(n_agents, 256), 256>>>(d_np, d_ni, n_agents, n_bld, d_bad);
        e = cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost);
    }
    if (e == cudaSuccess && !bad) {
        k_routines<<<grid_for(n_agents, 128), 128>>>(n_agents, V, k, d_home, d_anch, d_op, d_oi, d_ip, d_ii, d_np, d_ni, seed, d_r);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpy(routine, d_r, 4 * (size_t)n_agents * V, cudaMemcpyDeviceToHost);
    }
    for (void* p : tmp) cudaFree(p);
    if (e != cudaSuccess) return fail(nullptr, DYS_ERR_CUDA, "dys_generate_routines: %s", cudaGetErrorString(e));
    if (bad) return fail(nullptr, DYS_ERR_RANGE, "dys_generate_routines: invalid lists (flags 0x%x: 1 pointers not monotone, 2 index out of range, 4 list not strictly ascending)", bad);
    return DYS_OK;
}
