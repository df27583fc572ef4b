/* dysturbd.h -- C ABI of the B200-native per-day epidemic step of DySTUrbD.
 *
 * The reference (yair-grinb/DySTUrbD) is pure Python and has no FFI; the boundary below is reconstructed from
 * the function set of `epidemiological model.py` (update_buildings_status :8, update_counts :23,
 * calculate_hospitalizations :58, calculate_mortality :85, calculate_infections :108), the call order of
 * run_EM() (epidemiological_model.py:224-234) and the authoritative inline loop contagious3.py:176-387.
 * One dys_step() == one iteration of that `while` body (phases A-G of SURVEY.md section 3.2); phase H
 * (building_lockdown, contagious3.py:57-107) stays on the host and feeds dys_set_open().
 *
 * Conventions
 *   - plain C, no exceptions; every call returns DYS_OK (0) or a negative dys_status; dys_last_error(h)
 *     gives the message (h may be NULL for errors of dys_create).
 *   - one handle == one CUDA device + one stream; NOT thread-safe (one host thread per handle; replicas and
 *     shards use separate handles).
 *   - every array pointer may be a HOST pointer or a DEVICE pointer on the handle's device (e.g. a torch
 *     tensor's data_ptr()); the library copies on its own stream.  Caller keeps ownership of everything.
 *   - status codes are the reference's agents[:,13] values times two (contagious3.py:115-118):
 *       1 susceptible -> 2, 2 infected undiagnosed -> 4, 3 quarantined healthy -> 6, 3.5 quarantined infected
 *       -> 7, 4 diagnosed/home isolation -> 8, 5 hospitalised -> 10, 6 recovered -> 12, 7 dead -> 14.
 *   - agents, households, buildings, statistical areas are dense 0-based indices; the ORIGINAL row order of
 *     agents[] is significant (infector = largest infecting row index, contagious3.py:248; random keys).
 */
#ifndef DYSTURBD_H
#define DYSTURBD_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dys_handle dys_handle;

typedef enum {
    DYS_OK = 0,
    DYS_ERR_INVALID_ARG = -1,
    DYS_ERR_CUDA = -2,
    DYS_ERR_STATE = -3,     /* call order (e.g. step before load) */
    DYS_ERR_NOMEM = -4,
    DYS_ERR_RANGE = -5,     /* index out of range in an input array */
    DYS_ERR_NOT_AVAILABLE = -6,
    DYS_ERR_ABORTED = -7    /* a launch was abandoned: dys_abort(), or a grid barrier / peer wait timed out (DYS_SPIN_TIMEOUT_MS) */
} dys_status;

enum { DYS_S = 2, DYS_I = 4, DYS_QH = 6, DYS_QI = 7, DYS_D = 8, DYS_H = 10, DYS_R = 12, DYS_X = 14 };

/* dys_params.flags */
enum {
    DYS_WANT_SA_HIST = 1,        /* per-area days-since-infection histograms (compute_R per area, contagious3.py:282-291) */
    DYS_WANT_BUILDING_COUNTS = 2,/* infected residents per home building (contagious3.py:339-346) */
    DYS_WANT_EVENTS = 4,         /* today's (infected, infector) pairs ('Infections chain' / IO_mat inputs, :356-366) */
    DYS_WANT_IO_MAT = 8,         /* dense S x S origin matrix on the device (small S only) */
    DYS_TIME_KERNELS = 16        /* record CUDA events around each kernel (dys_get_kernel_times) */
};

#define DYS_PDF_LEN 64
#define DYS_HIST_LEN 21          /* days-since-infection 0..20: [0] = new infections, [1..recover-1] feed compute_R (:29-42) */
#define DYS_N_STATS 64

/* index into the int64 stats block returned by dys_get_stats() */
enum {
    DYS_ST_ACTIVE = 0,       /* #status in {2,3.5,4,5}                       'Active infected' (contagious3.py:326) */
    DYS_ST_DAILY_INF,        /* #status not in {1,3} and t_inf == day        'Daily infections' */
    DYS_ST_RECOVERED,        /* #6 */
    DYS_ST_QUARANTINED,      /* #3 <= status <= 4 */
    DYS_ST_DAILY_QUAR,       /* ... and t_q == day */
    DYS_ST_HOSPITALIZED,     /* #5 */
    DYS_ST_DAILY_HOSP,       /* admissions drawn today (:215) */
    DYS_ST_DEAD,             /* #7 */
    DYS_ST_DAILY_DEATHS,     /* deaths drawn today (:226) */
    DYS_ST_EVER_INFECTED,    /* #status not in {1,3} */
    DYS_ST_NEW_DIAG,         /* #status 4 with days-since-infection == diagnosis (today's diagnoses, :261) */
    DYS_ST_N_EVENTS,         /* infections applied today (phase D) */
    /* work counters of the day (feed the algorithmic-bytes formula of DESIGN.md) */
    DYS_ST_SUS0,             /* susceptible at day start (status 1 or 3) */
    DYS_ST_INF0,             /* infectious at day start (status 2, 3.5, 4) */
    /* 14..17 are diagnostics of the push.  Only non-zeros whose pair can ever share a building are kept (the others have
     * exposure 0 on every day).  dys_step: 16 and 17 equal the reference's count of P > 0 pairs and of P > U successes.
     * Inside a dys_step_many window the push of day d+1 runs before day d is finished everywhere and does not read the
     * row agent's status, so all four count pairs regardless of whether the row agent is still susceptible. */
    DYS_ST_EDGES_SCANNED,    /* kept non-zeros in the columns of today's infectious agents */
    DYS_ST_CANDIDATES,       /* ... whose row agent is susceptible */
    DYS_ST_EXPOSED,          /* ... and shares a building today: one Bernoulli each */
    DYS_ST_HITS,             /* Bernoulli successes */
    DYS_ST_CHANGED,          /* agents whose state was rewritten today */
    DYS_ST_COUNT_S = 20,     /* [20..27]: end-of-day agents in status 1, 2, 3, 3.5, 4, 5, 6, 7 */
    DYS_ST_CLOSED = 28,      /* device lockdown policy only: closed buildings of the day, R and Known R as IEEE double bits */
    DYS_ST_R_BITS = 29,
    DYS_ST_KNOWN_R_BITS = 30,
    DYS_ST_HIST0 = 32        /* [32 .. 32+20]: global days-since-infection histogram */
};

typedef struct dys_params {
    int64_t n_agents;            /* N (<= 2^27) */
    int64_t n_households;        /* H */
    int64_t n_buildings;         /* B */
    int64_t n_areas;             /* S */
    int32_t n_slots;             /* V: routine slots per agent (<= 16; 10 in the reference) */
    int32_t recover;             /* parameters.py:42  (21) */
    int32_t hospital_recover;    /* parameters.py:43  (28) */
    int32_t diagnosis;           /* parameters.py:45  (7)  */
    int32_t quarantine;          /* parameters.py:46  (7)  */
    int32_t adm_min_day;         /* contagious3.py:205 (4)  -- absolute infection day window for admission */
    int32_t adm_max_day;         /* contagious3.py:205 (14) */
    int32_t death_min_adm_day;   /* contagious3.py:219 (3)  */
    double norm_factor;          /* parameters.py:41  (0.08) */
    double compliance;           /* extension: P(household member obeys quarantine); 1.0 == reference */
    double pdf_table[DYS_PDF_LEN]; /* contagious_risk_day.pdf(0..63), host-computed (parameters.py:44) */
    uint64_t seed;               /* Philox key; counter = (a, b, day, stream) */
    int32_t flags;
    int32_t n_replicas;          /* Monte-Carlo replicas held by this handle (0 or 1 = a single run).  Replicas share the city
                                  * (network, routines, households) and differ in state, seed, norm_factor and compliance;
                                  * dys_step_many advances ALL of them in one launch.  Not with world > 1 / DYS_WANT_IO_MAT. */
    /* sharded mode (world > 1): this handle owns agents [shard_lo, shard_hi) of the n_agents global agents */
    int32_t rank, world;
    int64_t shard_lo, shard_hi;
    /* `routine` passed to dys_load_population covers global agents [halo_lo, halo_hi) -- the owned range plus every
     * agent an owned interaction_prob row references; 0, 0 means all n_agents */
    int64_t halo_lo, halo_hi;
    /* the households / home buildings / statistical areas of the owned agents form these index ranges (0, 0 = all).
     * hh / home / sa are passed as global indices; the per-household flags and the per-area / per-building OUTPUTS
     * (dys_get_sa_hist, dys_get_building_counts, dys_get_outputs) only cover the owned ranges, so a shard's daily
     * transfer does not grow with the total population.  hh_always / trig_hh of dys_load_quirk are relative to hh_lo. */
    int64_t hh_lo, hh_hi, bld_lo, bld_hi, area_lo, area_hi;
} dys_params;

/* literal pre-drawn uniforms (small N only): replaces np.random.random at contagious3.py:212, :224, :239 */
typedef struct dys_injected {
    const double* u_adm;         /* [N]   */
    const double* u_death;       /* [N]   */
    const double* u_pair;        /* [N*N] row = susceptible, col = infectious, original indices */
} dys_injected;

int dys_create(const dys_params* params, int device, dys_handle** out);
int dys_destroy(dys_handle* h);
const char* dys_last_error(const dys_handle* h);

/* agents[N,29] as structure-of-arrays (host wrapper: dysturbd_b200/population.py).
 * status_x2 <- col 13, t_inf <- col 16, t_q <- col 19, t_adm <- col 24, src <- row index of col 21 (-1 none),
 * hh <- col 1, home <- col 7, sa <- col 22 (-1 = none), risk <- col 14, p_adm <- col 23, p_death <- col 26,
 * routine <- bld_visits_by_agents[N,V] as building indices, -1 = empty (contagious3.py:139-170). */
int dys_load_population(dys_handle* h, const uint8_t* status_x2, const int32_t* t_inf, const int32_t* t_q,
                        const int32_t* t_adm, const int32_t* src, const int32_t* hh, const int32_t* home,
                        const int32_t* sa, const double* risk, const double* p_adm, const double* p_death,
                        const int32_t* routine);
/* interaction_prob (contagious3.py:129) as CSR over its non-zeros; columns ascending within a row.
 * Must follow dys_load_population (the per-edge shared-building classes are derived from the routines). */
int dys_load_graph(dys_handle* h, const int64_t* rowptr, const int32_t* col, const double* val);
/* the whole-row isin of contagious3.py:267; all three may be NULL (no quirk) */
int dys_load_quirk(dys_handle* h, const uint8_t* hh_always, const int64_t* trig_ptr, const int32_t* trig_hh);
/* build[:,10] for the coming day (contagious3.py:181, 384) */
int dys_set_open(dys_handle* h, const uint8_t* open_flags);
/* parameter sweeps / Monte-Carlo replicas on a loaded population */
int dys_set_run(dys_handle* h, double norm_factor, double compliance, uint64_t seed);
/* n_replicas > 1: the per-replica run parameters (arrays of n_replicas; NULL = keep).  dys_load_population / dys_set_state
 * give every replica the same initial state; dys_select_replica(r) points the getters (dys_get_outputs, dys_get_stats,
 * dys_get_sa_hist, dys_get_building_counts, dys_get_events, dys_get_state) and dys_set_state at replica r
 * (r = -1: dys_set_state writes all replicas, getters read replica 0). */
int dys_set_replica_params(dys_handle* h, const double* norm_factor, const double* compliance, const uint64_t* seed);
int dys_select_replica(dys_handle* h, int32_t r);
/* Phase H on the device (contagious3.py:57-107, 368-385).  scenario = OR of DYS_SC_* (0 switches the device policy off
 * again); bld_class[B] = OR of DYS_CLASS_NONRES (build[:,1] >= 3), DYS_CLASS_EDU, DYS_CLASS_REL (build[:,9] in the code lists
 * of :67-74).  From then on dys_step_many decides every day's open flags on the device from the lagged R (open_flags must be
 * NULL) and reports 'Closed buildings', R and Known R (IEEE double bits) in stats[28..30].  Exact for the deterministic
 * scenarios (ALL / EDU / REL); DYS_SC_GRADUAL closes each building of the class with probability 1/2 on a keyed draw where
 * the reference draws exactly half with np.random.choice (keep the host policy + dys_step_many(open_flags) for that).
 * Call after dys_load_population / dys_set_state, before day 0; needs 2 <= diagnosis <= 12; not in sharded mode. */
enum { DYS_SC_GRADUAL = 1, DYS_SC_ALL = 2, DYS_SC_EDU = 4, DYS_SC_REL = 8 };
enum { DYS_CLASS_NONRES = 2, DYS_CLASS_EDU = 4, DYS_CLASS_REL = 8 };
int dys_set_lockdown(dys_handle* h, int32_t scenario, const uint8_t* bld_class);
/* The DIFF scenarios (contagious3.py:368-376): the same rule applied area by area -- bld_area[B] = the building's area
 * (row of the per-area histogram, 0 <= bld_area[b] < S; build[:,3] ranked), each area judged on the R of its own residents
 * `diagnosis` days back (sas_vis_R, :311-318) and the day before.  As shipped, the reference cannot run these scenarios: :371
 * reads SAs[day-1]['Vis_R'] while :319 writes 'vis_R' (KeyError on day 1); this entry point implements the loop with
 * that key read as written (pinned by the tests/golden '*_keyfix' fixtures, made from the reference's text with the one
 * key repaired).  Needs DYS_WANT_SA_HIST; the other conditions of dys_set_lockdown apply. */
int dys_set_lockdown_areas(dys_handle* h, int32_t scenario, const uint8_t* bld_class, const int32_t* bld_area);
/* asks a running launch to end at its next spin check (grid barrier / peer wait); the handle then reports
 * DYS_ERR_ABORTED until dys_set_state / dys_load_population.  Safe to call from another host thread. */
int dys_abort(dys_handle* h);
int dys_set_state(dys_handle* h, const uint8_t* status_x2, const int32_t* t_inf, const int32_t* t_q,
                  const int32_t* t_adm, const int32_t* src);

/* one day: contagious3.py:176-366, asynchronous on the handle's stream */
int dys_step(dys_handle* h, int32_t day, const dys_injected* injected /* nullable */);
/* the same day in two halves, for sharded populations: _begin writes this shard's slice of the day-start
 * infectious bytes (DYS_BUF_INFECTIOUS) and applies phases B, C; the caller exchanges the slices between ranks
 * (all-gather over NCCL); _end runs phases D-G.  dys_step == dys_step_begin + dys_step_end. */
/* `n_days` consecutive days queued in one call.  open_flags is NULL (flags unchanged) or [n_days][B], one vector per
 * day.  Exact for every scenario of the reference as long as n_days <= `diagnosis` (7): the lockdown decision for a day
 * uses R of `diagnosis` days earlier (contagious3.py:306-309, 379-384), so a week's flags are known when it starts. */
int dys_step_many(dys_handle* h, int32_t first_day, int32_t n_days, const uint8_t* open_flags /* nullable */);
int dys_step_begin(dys_handle* h, int32_t day, const dys_injected* injected /* nullable */);
int dys_step_end(dys_handle* h, int32_t day);

/* per-day integer outputs; these wait for `day` to finish.  Stats of the last DYS_STATS_RING days are kept;
 * histograms, building counts, events and the IO matrix only for the most recent day. */
#define DYS_STATS_RING 64
int dys_get_stats(dys_handle* h, int32_t day, int64_t* out /* [DYS_N_STATS] */);
int dys_get_sa_hist(dys_handle* h, int32_t day, int32_t* out /* [(area_hi-area_lo)*DYS_HIST_LEN] */);
int dys_get_building_counts(dys_handle* h, int32_t day, int32_t* out /* [bld_hi-bld_lo] */);
int dys_get_events(dys_handle* h, int32_t day, int32_t* agent, int32_t* infector, int64_t cap, int64_t* n);
int dys_get_io_mat(dys_handle* h, int32_t day, int32_t* out /* [S*S] */);
int dys_get_state(dys_handle* h, uint8_t* status_x2, int32_t* t_inf, int32_t* t_q, int32_t* t_adm, int32_t* src);
/* Everything a day produced, in ONE device->host transfer that dys_step() already queued into library-owned pinned host
 * memory: this call only waits for it.  The pointers stay valid until DYS_OUT_RING - 1 further days have been stepped,
 * so a host loop can queue day d+1 before it consumes day d (the lockdown decision lags `diagnosis` = 7 days,
 * contagious3.py:306-309, so nothing a day produces is needed to start the next one). */
#define DYS_OUT_RING 16
typedef struct dys_day_outputs {
    int32_t day;
    int32_t events_inline;       /* how many of the n_events pairs are in `events` (the rest: dys_get_events) */
    int64_t n_events;
    const int64_t* stats;        /* [DYS_N_STATS] */
    const int32_t* sa_hist;      /* [S*DYS_HIST_LEN] or NULL */
    const int32_t* building_counts; /* [B] or NULL */
    const int32_t* events;       /* [events_inline][2] = (infected, infector) row indices, or NULL */
    int64_t events_capacity;     /* pairs the block can carry: events[0 .. events_capacity) is addressable on every day */
} dys_day_outputs;
int dys_get_outputs(dys_handle* h, int32_t day, dys_day_outputs* out);
/* Everything the host loop needs from one day in ONE call: waits for the day's block (like dys_get_outputs), copies stats,
 * per-building counts and the (infected, infector) pairs out of the pinned ring into caller memory, and forms R
 * (compute_R, contagious3.py:29-42, same order of float operations) for the global histogram and for every owned area. */
int dys_consume_day(dys_handle* h, int32_t day, const double* pdf, int32_t recover, int64_t* stats /* [DYS_N_STATS] */,
                    double* R_global /* [1] */, double* R_areas /* [areas] or NULL */, int32_t* building_counts /* [buildings] or NULL */,
                    int32_t* events /* [events_cap][2] or NULL */, int64_t events_cap, int64_t* n_events /* nullable */);
/* P for every evaluated pair of the most recent day (tests; requires injected mode): out[N*N] */
int dys_get_pair_prob(dys_handle* h, double* out);

int dys_sync(dys_handle* h);
/* mean milliseconds per launch and launch counts since the last call, per kernel
 * (0 k_snapshot -- only launched on the first day after a load, 1 k_push, 2 k_agent); needs DYS_TIME_KERNELS */
#define DYS_N_KERNELS 3
int dys_get_kernel_times(dys_handle* h, double* ms_total /* [DYS_N_KERNELS] */, int64_t* launches /* [DYS_N_KERNELS] */);
/* raw device pointers for zero-copy interop (torch / DLPack / NCCL): which = one of dys_buffer */
/* DYS_BUF_AGENT_REC: [n_pad] 16-byte records {int16 t_inf, t_q, t_adm, pad; int32 area (local), home building (local)} */
typedef enum { DYS_BUF_STATUS = 0, DYS_BUF_INFECTIOUS = 1, DYS_BUF_STATS = 2, DYS_BUF_AGENT_REC = 3 } dys_buffer;
int dys_device_buffer(dys_handle* h, int which, void** ptr, int64_t* bytes);
int dys_cuda_stream(dys_handle* h, void** stream);
int64_t dys_device_bytes(const dys_handle* h);
/* compute_R (contagious3.py:29-42) on `rows` histograms held in HOST memory (int32 or int64 elements, DYS_HIST_LEN used
 * per row, rows `row_stride` elements apart): out[r] = hist[r][0] / sum_{i=1}^{recover-1} hist[r][i] * pdf[i], 0 if the sum
 * is 0 -- the reference's order of float operations, so the results compare with ==. */
int dys_compute_r_rows(const void* hist, int32_t elem_bytes, int64_t rows, int64_t row_stride, const double* pdf,
                       int32_t recover, double* out);
/* bytes of open flags uploaded host -> device by dys_step_many so far (benchmark accounting) */
int64_t dys_h2d_bytes(const dys_handle* h);
/* Peer mode (sharded populations on one NVLink/NVSwitch node): every rank exports CUDA IPC handles of its snapshot
 * buffer and flag array, gathers all ranks' handles (any host-side all-gather) and attaches them.  From then on the
 * kernel that produces tomorrow's snapshot bytes stores them straight into every peer GPU's buffer and raises a flag
 * there; the next day's first kernel waits for all flags.  dys_step() then needs no collective call at all.  After
 * dys_set_state in peer mode the caller must barrier across ranks before the next dys_step. */
#define DYS_PEER_HANDLE_BYTES 128
int dys_peer_export(dys_handle* h, void* out /* [DYS_PEER_HANDLE_BYTES] */);
int dys_peer_detach(dys_handle* h);   /* back to the collective exchange (all ranks must do the same) */
int dys_peer_attach(dys_handle* h, int n_ranks, const void* handles /* [n_ranks][DYS_PEER_HANDLE_BYTES] */,
                    const int64_t* ranges /* [n_ranks][4]: every rank's shard_lo, shard_hi, halo_lo, halo_hi -- only the
                                             bytes a peer reads travel, and only the ranks that exchange bytes synchronise */);
/* evaluates the keyed uniform u(day, stream, a[k], b[k]) ON THE DEVICE for n host-side index pairs (known-answer
 * tests of the device Philox against oracle/philox_np.py) */
int dys_keyed_uniform_device(int device, uint64_t seed, uint32_t day, uint32_t stream, const uint32_t* a,
                             const uint32_t* b, int64_t n, double* out);
/* Routine generation (contagious3.py:139-170) on the device: every agent's static visit list [n_agents][V], -1 padded --
 * home, then before each anchor (anchors[a][0..2] <- agents[:,12], agents[:,18], home; -1 = NaN) up to k flexible stops, each
 * taken on a coin flip (:150) and drawn uniformly (:155) from (in[anchor] ∩ out[current position]) ∪ near[agent] in ascending
 * building order.  Buildings are numbered by ASCENDING ID (the order np.union1d enumerates them in); out[b] / in[b] =
 * buildings within bld_dist FROM / TO b, near[a] = buildings within a_dist of agent a, as CSR lists with ascending indices
 * (dysturbd_b200/routines.py derives them from the reference's shortest-path matrix).  Draws are keyed:
 * coin = u(seed, 0, 5, a, q) >= 0.5, choice = (int)(u(seed, 0, 6, a, q) * len), q = index of the coin among the agent's coins.
 * All pointers are HOST pointers; V >= 1 + 3 (k + 1). */
int dys_generate_routines(int device, int64_t n_agents, int64_t n_bld, int32_t V, int32_t k, const int32_t* home,
                          const int32_t* anchors, const int64_t* out_ptr, const int32_t* out_idx, const int64_t* in_ptr,
                          const int32_t* in_idx, const int64_t* near_ptr, const int32_t* near_idx, uint64_t seed,
                          int32_t* routine);
/* Network construction (communities.py:22-40, 65-70 and contagious3.py:124-130) -- the two O(N^2) steps before the day loop.
 * All pointers are HOST pointers.
 *
 * dys_similarity_edges: the agent-agent edges of create_network without its N x N matrices.  Per agent: age (agents[:,4]),
 * household income (households[agent_hh, 2]), home coordinates (build[agent_homes, 6:8]) and household row (agent_hh).
 * For every ordered pair i != j,  p = ((((w_a * fa) * w_i) * fi) * w_d) * fd  with  f = 1 - |difference| / max |difference|
 * (Euclidean home distance for fd), p = 1 inside a household; an edge wherever p > min_prob.  Same float64 operations in the
 * same order as the reference, so `prob` is bit-identical to agents_prob[i, j]; the caller forms the weight -log(p).
 * Output: CSR with rows and columns ascending (np.where order).  indptr[n+1] is always written; with col == prob == NULL the
 * call only sizes (indptr[n] = number of edges), otherwise capacity must hold them.  maxes[3] (nullable) = the normalisers. */
int dys_similarity_edges(int device, int64_t n, const double* age, const double* income, const double* x, const double* y,
                         const int64_t* household, double w_a, double w_i, double w_d, double min_prob, int64_t* indptr,
                         int64_t capacity, int32_t* col, double* prob, double* maxes);
/* dys_building_edges: the building-building edges of create_network (communities.py:12-21, 53-54), the gravity model
 *   scores[i][j] = floor_space[j] / dist(i, j)^2   (the square of the ROUNDED Euclidean distance; 0 -> 0.00001; scores[i][i] = 0)
 *   prob = scores / row sum,   an edge i -> j wherever prob > min_prob, i != j.
 * The row sum is numpy's pairwise `scores.sum(axis=1)`, reproduced addition by addition, so `prob` is bit-identical to
 * bld_prob[i, j]; the caller forms -log(prob).  Output protocol as dys_similarity_edges. */
int dys_building_edges(int device, int64_t n_bld, const double* floor_space, const double* x, const double* y, double min_prob,
                       int64_t* indptr, int64_t capacity, int32_t* col, double* prob);
/* dys_shortest_paths: scipy.sparse.csgraph.shortest_path(g, directed=True, indices=sources) for a CSR graph with non-negative
 * weights (explicit zeros are edges, as in scipy): out[k][v] = length of the shortest path sources[k] -> v, +inf where there
 * is none.  Bit-identical to Dijkstra's left-to-right sums.  Give the sources grouped by connected component:
 * a batch of sources only touches the vertices its sources can reach.  info[4] (nullable) = relaxation sweeps that changed
 * something (summed over the batches), then seconds spent on preparation + upload, in the relaxation, on transpose + download. */
int dys_shortest_paths(int device, int64_t n, const int64_t* indptr, const int32_t* indices, const double* weights,
                       int64_t n_sources, const int32_t* sources, double* out, double* info);
/* the same distances, but only the vertices each source reaches, compacted on the device: out_ptr[n_sources + 1], then
 * (out_col, out_dist) with the vertices of a source ascending -- a third of the bytes of the dense rows on the shipped city.
 * skip_self != 0 leaves the source itself out (contagious3.py:126 sets the diagonal to inf before :127 takes exp(-dists)).
 * max_dist: only vertices with distance < max_dist are listed (INFINITY: every reachable vertex) -- and the search itself stops
 * there (a label at or beyond the bound is never stored), which is exact for every listed entry: this is the bounded search
 * behind the routine generation's neighbourhoods, `dists[...] < bld_dist` / `< a_dist` (contagious3.py:144, 148, 152).
 * col_keep[n] (nullable): list only the marked vertices (e.g. the building nodes).
 * capacity = entries out_col / out_dist can hold (the sizes of the sources' weak components bound it); DYS_ERR_RANGE if short. */
int dys_shortest_paths_csr(int device, int64_t n, const int64_t* indptr, const int32_t* indices, const double* weights,
                           int64_t n_sources, const int32_t* sources, int32_t skip_self, double max_dist, const uint8_t* col_keep,
                           int64_t* out_ptr, int64_t capacity, int32_t* out_col, double* out_dist, double* info);
const char* dys_version(void);

#ifdef __cplusplus
}
#endif
#endif
