/* TEST INFRASTRUCTURE (oracle) -- plain-C CPU restatement of one day of DySTUrbD's epidemic step.
 *
 * NOT the product.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product path (dysturbd_b200/) never does and has no CPU fallback.
 *
 * What it restates: /root/reference/contagious3.py:176-366 (the per-day `while` body), phase by phase, on a
 * structure-of-arrays state instead of the reference's agents[N,29] float64 matrix, every per-agent loop an OpenMP
 * loop (so that the CPU baseline of bench.py really uses the host cores it reports), and with the dense
 * U x I pair matrices of :230-248 replaced by a walk over the non-zeros of interaction_prob (a pair with
 * interaction_prob == 0 has P == 0 and `0 > u` is never true, so the sparse walk is exact).
 * Parity status: PINNED against the literal reference loop run under identical injected draws
 * (oracle/literal.py -> tests/golden/ fixtures, tests/test_oracle_vs_literal.py).  The reference itself
 * ships no tests or golden vectors.
 *
 * Status codes are the reference's float codes times two (1,2,3,3.5,4,5,6,7 -> 2,4,6,7,8,10,12,14).
 * Build:  gcc -O2 -fopenmp -shared -fPIC -ffp-contract=off oracle/dys_oracle.c -o oracle/libdys_oracle.so
 *         (-ffp-contract=off: the probability chain must be four separately rounded IEEE multiplies.)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ST_S 2
#define ST_I 4
#define ST_QH 6
#define ST_QI 7
#define ST_D 8
#define ST_H 10
#define ST_R 12
#define ST_X 14

/* ---- Philox4x32-10 (Salmon et al., SC'11; Random123 constants), see oracle/philox_np.py ---- */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr, key, out); }

double orc_keyed_uniform(uint64_t seed, uint32_t day, uint32_t stream, uint32_t a, uint32_t b)
{
    uint32_t ctr[4] = {a, b, day, stream}, key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, o[4];
    philox4x32_10(ctr, key, o);
    uint64_t w = (uint64_t)o[0] | ((uint64_t)o[1] << 32);
    return (double)(w >> 11) * (1.0 / 9007199254740992.0);
}

void orc_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- the day ---- */
typedef struct {
    int64_t N, H, B, S;
    int32_t V;
    int32_t recover, hospital_recover, diagnosis, quarantine; /* parameters.py:42-46 */
    int32_t adm_min_day, adm_max_day, death_min_adm_day;       /* contagious3.py:205 (4, 14), :219 (3) */
    double norm_factor;                                        /* parameters.py:41 */
    double compliance;                                         /* extension; 1.0 == reference */
    double pdf[64];                                            /* contagious_risk_day.pdf(0..63), parameters.py:44 */
    uint64_t seed;
    /* state (mutated) */
    uint8_t* status;
    int32_t *t_inf, *t_q, *t_adm, *src;
    /* static per-agent */
    const int32_t *hh, *home, *sa;
    const double *risk, *p_adm, *p_death;
    const int32_t* routine;          /* [N*V] building index or -1 */
    /* interaction_prob CSR */
    const int64_t* rowptr;
    const int32_t* col;
    const double* val;
    /* buildings */
    const uint8_t* open;             /* [B] */
    /* household-isin quirk of contagious3.py:267 */
    const uint8_t* hh_always;        /* [H] or NULL */
    const int64_t* trig_ptr;         /* [N+1] or NULL */
    const int32_t* trig_hh;
    /* injected draws (NULL -> keyed Philox) */
    const double *u_adm, *u_death, *u_pair;   /* [N], [N], [N*N] */
    /* outputs */
    int64_t* stats;                  /* [64]: counters 0..31, global days-since-infection histogram 32..52 */
    int32_t* sa_hist;                /* [S*21] or NULL */
    int32_t* bld_infected;           /* [B] or NULL */
    int32_t *ev_agent, *ev_src;      /* [N] or NULL: today's infections in ascending agent order */
    double* pair_prob;               /* [N*N] or NULL: P written for every evaluated (u,i) pair (tests) */
    /* sharded populations: this context owns global agents [lo, lo+N) of n_glob; `routine` covers global agents from
     * rt_lo; interaction_prob columns, src, the random keys and the events are GLOBAL indices.  ib_global (when not
     * NULL) is the all-gathered day-start infectious byte of every agent: bit7 infectious (status 2/3.5/4), bit6
     * isolated (3.5/4), bits0-5 days since infection -- the only thing a day needs to know about a remote agent. */
    int64_t lo, n_glob, rt_lo;
    const uint8_t* ib_global;
} orc_ctx;

enum { O_ACTIVE, O_DAILY_INF, O_RECOVERED, O_QUARANTINED, O_DAILY_QUAR, O_HOSP, O_DAILY_HOSP, O_DEAD, O_DAILY_DEATHS,
       O_EVER_INFECTED, O_NEW_DIAG, O_N_EVENTS, O_SUS0, O_INF0, O_EDGES_SCANNED, O_CANDIDATES, O_EXPOSED, O_HITS };

static uint8_t snapshot_byte(uint8_t st, int day, int t_inf)
{
    if (!(st == ST_I || st == ST_QI || st == ST_D)) return 0;
    int n = day - t_inf;
    if (n < 0) n = 0;
    if (n > 63) n = 63;
    return (uint8_t)(0x80 | (st != ST_I ? 0x40 : 0) | n);
}

/* day-start infectious bytes of the owned agents (what a shard contributes to the per-day exchange) */
void orc_snapshot(const orc_ctx* c, int32_t day, uint8_t* out)
{
    for (int64_t a = 0; a < c->N; ++a) out[a] = snapshot_byte(c->status[a], day, c->t_inf[a]);
}

/* a = row of the routine table (global agent index - rt_lo); gone = admitted or dead; iso = keeps only the home slot */
static int eff_list(const orc_ctx* c, int64_t a, int gone, int iso, int32_t* out)
{
    /* contagious3.py:181-189: closed buildings removed; isolated agents keep slot 0; admitted/dead keep nothing */
    int n = 0;
    if (gone) return 0;
    for (int s = 0; s < c->V; ++s) {
        int32_t b = c->routine[a * c->V + s];
        if (b < 0) continue;
        if (!c->open[b]) continue;
        if (s > 0 && iso) continue;
        out[n++] = b;
    }
    return n;
}

int orc_day(orc_ctx* c, int32_t day)
{
    const int64_t N = c->N;
    uint8_t* st0 = (uint8_t*)malloc(N);
    if (!st0) return -1;
    memcpy(st0, c->status, N);
    memset(c->stats, 0, 64 * sizeof(int64_t));
    int64_t n_adm = 0, n_death = 0, n_sus0 = 0, n_inf0 = 0;

    /* phases B and C (contagious3.py:205-227), both on the day-start snapshot.  (Every loop over agents is an OpenMP loop:
     * agents are independent inside a phase, so the thread count changes nothing but the time.) */
#pragma omp parallel for schedule(static) reduction(+ : n_adm, n_death, n_sus0, n_inf0)
    for (int64_t a = 0; a < N; ++a) {
        uint8_t s = st0[a];
        int inf0 = (s == ST_I || s == ST_QI || s == ST_D);
        n_inf0 += inf0;
        n_sus0 += (s == ST_S || s == ST_QH);
        if (inf0 && c->t_inf[a] >= c->adm_min_day && c->t_inf[a] <= c->adm_max_day) {
            double u = c->u_adm ? c->u_adm[c->lo + a] : orc_keyed_uniform(c->seed, day, 0, (uint32_t)(c->lo + a), 0);
            if (c->p_adm[a] > u) { c->status[a] = ST_H; c->t_adm[a] = day; ++n_adm; }
        }
        if (s == ST_H && c->t_adm[a] >= c->death_min_adm_day) {
            double u = c->u_death ? c->u_death[c->lo + a] : orc_keyed_uniform(c->seed, day, 1, (uint32_t)(c->lo + a), 0);
            if (c->p_death[a] > u) { c->status[a] = ST_X; ++n_death; }
        }
    }

    /* phase D (contagious3.py:230-248) */
    int64_t scanned = 0, cand = 0, expo = 0, hits = 0;
#pragma omp parallel for schedule(dynamic, 512) reduction(+ : scanned, cand, expo, hits)
    for (int64_t u = 0; u < N; ++u) {
        uint8_t su = st0[u];
        if (!(su == ST_S || su == ST_QH)) continue;
        int32_t eu[16], ei[16];
        const int64_t ug = c->lo + u;
        int nu = eff_list(c, ug - c->rt_lo, 0, su == ST_QH, eu);
        int32_t best = -1;
        for (int64_t e = c->rowptr[u]; e < c->rowptr[u + 1]; ++e) {
            int32_t i = c->col[e];                       /* GLOBAL index of the column agent */
            ++scanned;
            int inf_i, iso_i, n_inf;
            if (c->ib_global) {
                uint8_t b = c->ib_global[i];
                inf_i = b & 0x80; iso_i = (b & 0x40) != 0; n_inf = b & 63;
            } else {
                uint8_t si = st0[i];
                inf_i = (si == ST_I || si == ST_QI || si == ST_D);
                iso_i = (si == ST_QI || si == ST_D);
                n_inf = day - c->t_inf[i];
                if (n_inf > 63) n_inf = 63;
            }
            if (!inf_i) continue;
            ++cand;
            int ni = eff_list(c, (int64_t)i - c->rt_lo, 0, iso_i, ei), share = 0;
            for (int x = 0; x < nu && !share; ++x)
                for (int y = 0; y < ni; ++y)
                    if (eu[x] == ei[y]) { share = 1; break; }
            double p = c->val[e] * c->risk[u];          /* :232-233 */
            p = p * c->pdf[n_inf];                      /* :234 */
            p = p * (share ? 1.0 : 0.0);                /* :237 */
            p = p * c->norm_factor;                     /* :238 */
            if (c->pair_prob) c->pair_prob[ug * c->n_glob + i] = p;
            if (!share) continue;
            ++expo;
            double r = c->u_pair ? c->u_pair[ug * c->n_glob + i] : orc_keyed_uniform(c->seed, day, 2, (uint32_t)ug, (uint32_t)i);
            if (p > r) { ++hits; if (i > best) best = i; }   /* :240, :248 last writer = largest index */
        }
        if (best >= 0) {
            c->status[u] = (su == ST_S) ? ST_I : ST_QI;  /* :243, :245 */
            c->t_inf[u] = day;
            c->src[u] = best;
        }
    }

    /* phase E (contagious3.py:250-259), rule by rule; day counters derived as day - start */
#pragma omp parallel for schedule(static)
    for (int64_t a = 0; a < N; ++a) {
        uint8_t s = c->status[a];
        int n_inf = day - c->t_inf[a], n_q = day - c->t_q[a];
        if (s == ST_QH && n_q == c->quarantine) s = ST_S;
        if (s == ST_QI && n_q == c->quarantine) s = ST_I;
        if (s == ST_I && n_inf == c->diagnosis) c->t_q[a] = day;
        if ((s == ST_I || s == ST_QI) && n_inf == c->diagnosis) s = ST_D;
        if (s == ST_D && n_inf == c->recover) s = ST_R;
        if (s == ST_H && n_inf == c->hospital_recover) s = ST_R;
        if (s == ST_S || s == ST_I || s == ST_R || s == ST_X) c->t_adm[a] = 0;
        c->status[a] = s;
    }

    /* phase F (contagious3.py:261-270) */
    int64_t n_diag = 0;
    uint8_t* hh_hit = (uint8_t*)calloc(c->H > 0 ? c->H : 1, 1);
    uint8_t* hh_quirk = (uint8_t*)calloc(c->H > 0 ? c->H : 1, 1);
#pragma omp parallel for schedule(static) reduction(+ : n_diag)
    for (int64_t a = 0; a < N; ++a) {
        if (c->status[a] == ST_D && day - c->t_inf[a] == c->diagnosis) {
            ++n_diag;
#pragma omp atomic write
            hh_hit[c->hh[a]] = 1;
            if (c->trig_ptr)
                for (int64_t t = c->trig_ptr[a]; t < c->trig_ptr[a + 1]; ++t) {
#pragma omp atomic write
                    hh_quirk[c->trig_hh[t]] = 1;
                }
        }
    }
    if (n_diag > 0) {
#pragma omp parallel for schedule(static)
        for (int64_t a = 0; a < N; ++a) {
            int32_t h = c->hh[a];
            uint8_t s = c->status[a];
            int obey = 1;
            if (c->compliance < 1.0) obey = orc_keyed_uniform(c->seed, day, 3, (uint32_t)(c->lo + a), 0) < c->compliance;
            if (s == ST_S && hh_hit[h] && obey) { c->status[a] = ST_QH; c->t_q[a] = day; }
            else if (s == ST_I && (hh_hit[h] || hh_quirk[h] || (c->hh_always && c->hh_always[h])) && obey) {
                c->status[a] = ST_QI; c->t_q[a] = day;
            }
        }
    }
    free(hh_hit); free(hh_quirk);

    /* phase G: integer statistics (contagious3.py:273-366); floats (R) are formed on the host */
    int64_t* S = c->stats;
    if (c->sa_hist) memset(c->sa_hist, 0, sizeof(int32_t) * 21 * c->S);
    if (c->bld_infected) memset(c->bld_infected, 0, sizeof(int32_t) * c->B);
    int64_t g_active = 0, g_dinf = 0, g_rec = 0, g_quar = 0, g_dquar = 0, g_hosp = 0, g_dead = 0, g_ever = 0;
    int64_t hist[21];
    memset(hist, 0, sizeof hist);
    /* today's infections are listed in ascending agent order: count per chunk, prefix, fill */
    enum { CHUNK = 1 << 15 };
    const int64_t n_chunks = (N + CHUNK - 1) / CHUNK;
    int64_t* chunk_ev = (int64_t*)calloc((size_t)n_chunks + 1, sizeof(int64_t));
#pragma omp parallel for schedule(static) reduction(+ : g_active, g_dinf, g_rec, g_quar, g_dquar, g_hosp, g_dead, g_ever, hist[:21])
    for (int64_t ch = 0; ch < n_chunks; ++ch) {
        int64_t nev_c = 0;
        const int64_t a1 = (ch + 1) * CHUNK < N ? (ch + 1) * CHUNK : N;
        for (int64_t a = ch * CHUNK; a < a1; ++a) {
            uint8_t s = c->status[a];
            int ever = !(s == ST_S || s == ST_QH);
            g_active += (s == ST_I || s == ST_QI || s == ST_D || s == ST_H);
            g_dinf += (ever && c->t_inf[a] == day);
            g_rec += (s == ST_R);
            g_quar += (s == ST_QH || s == ST_QI || s == ST_D);
            g_dquar += ((s == ST_QH || s == ST_QI || s == ST_D) && c->t_q[a] == day);
            g_hosp += (s == ST_H);
            g_dead += (s == ST_X);
            g_ever += ever;
            if (ever) {
                int k = day - c->t_inf[a];
                if (k >= 0 && k <= 20) {
                    hist[k] += 1;
                    if (c->sa_hist) {
#pragma omp atomic
                        c->sa_hist[c->sa[a] * 21 + k] += 1;
                    }
                }
            }
            if (c->bld_infected && (s == ST_I || s == ST_QI || s == ST_D)) {
#pragma omp atomic
                c->bld_infected[c->home[a]] += 1;
            }
            if (ever && c->t_inf[a] == day && st0[a] != s && (st0[a] == ST_S || st0[a] == ST_QH)) ++nev_c;
        }
        chunk_ev[ch + 1] = nev_c;
    }
    for (int64_t ch = 0; ch < n_chunks; ++ch) chunk_ev[ch + 1] += chunk_ev[ch];
    const int64_t nev = chunk_ev[n_chunks];
    if (c->ev_agent && nev > 0) {
#pragma omp parallel for schedule(static)
        for (int64_t ch = 0; ch < n_chunks; ++ch) {
            int64_t k = chunk_ev[ch];
            if (chunk_ev[ch + 1] == k) continue;
            const int64_t a1 = (ch + 1) * CHUNK < N ? (ch + 1) * CHUNK : N;
            for (int64_t a = ch * CHUNK; a < a1; ++a) {
                uint8_t s = c->status[a];
                int ever = !(s == ST_S || s == ST_QH);
                if (ever && c->t_inf[a] == day && st0[a] != s && (st0[a] == ST_S || st0[a] == ST_QH)) {
                    c->ev_agent[k] = (int32_t)(c->lo + a);
                    c->ev_src[k] = c->src[a];
                    ++k;
                }
            }
        }
    }
    free(chunk_ev);
    S[O_ACTIVE] = g_active; S[O_DAILY_INF] = g_dinf; S[O_RECOVERED] = g_rec; S[O_QUARANTINED] = g_quar;
    S[O_DAILY_QUAR] = g_dquar; S[O_HOSP] = g_hosp; S[O_DEAD] = g_dead; S[O_EVER_INFECTED] = g_ever;
    for (int k = 0; k < 21; ++k) S[32 + k] = hist[k];
    S[O_DAILY_HOSP] = n_adm;
    S[O_DAILY_DEATHS] = n_death;
    S[O_NEW_DIAG] = n_diag;
    S[O_N_EVENTS] = nev;
    S[O_SUS0] = n_sus0;
    S[O_INF0] = n_inf0;
    S[O_EDGES_SCANNED] = scanned;
    S[O_CANDIDATES] = cand;
    S[O_EXPOSED] = expo;
    S[O_HITS] = hits;
    free(st0);
    return 0;
}

/* ---- routine generation (contagious3.py:139-170), TEST INFRASTRUCTURE like the rest of this file ----
 * The reference walks, for every agent, its anchors (agents[:,12], agents[:,18], home) and before each one takes up to k
 * flexible stops: a coin (np.random.randint(2), :150), then a uniform choice (:155) from
 *     (buildings within bld_dist TO the anchor  ∩  buildings within bld_dist FROM the current position)  ∪
 *     (buildings within a_dist of the agent)
 * in ascending building-id order (np.intersect1d / np.union1d sort).  Here the three sets are sorted index lists
 * (buildings renumbered by ascending id) and the draws are keyed: coin = u(seed,0,5,agent,q) >= 0.5,
 * choice index = (int)(u(seed,0,6,agent,q) * len), q = index of the coin among the agent's coins -- what
 * oracle/literal.py injects into the unmodified reference code (tests/golden/routines_*.npz). */
static int64_t orc_union_pick(const int32_t* A, int64_t na, const int32_t* B, int64_t nb, const int32_t* C, int64_t nc,
                              int64_t pick /* -1: count only */, int32_t* out)
{
    int64_t pa = 0, pb = 0, pc = 0, count = 0;
    for (;;) {
        while (pa < na && pb < nb && A[pa] != B[pb]) { if (A[pa] < B[pb]) ++pa; else ++pb; }
        const int64_t x = (pa < na && pb < nb) ? A[pa] : INT64_MAX;
        const int64_t y = pc < nc ? C[pc] : INT64_MAX;
        const int64_t v = x < y ? x : y;
        if (v == INT64_MAX) break;
        if (count == pick) { *out = (int32_t)v; return count; }
        ++count;
        if (x == v) { ++pa; ++pb; }
        if (y == v) ++pc;
    }
    return count;
}

int orc_routines(int64_t N, int32_t V, int32_t k, const int32_t* home, const int32_t* anchors,
                 const int64_t* out_ptr, const int32_t* out_idx, const int64_t* in_ptr, const int32_t* in_idx,
                 const int64_t* near_ptr, const int32_t* near_idx, uint64_t seed, int32_t* routine)
{
    if (V < 1 + 3 * (k + 1)) return -1;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t a = 0; a < N; ++a) {
        int32_t* r = routine + a * V;
        for (int s = 0; s < V; ++s) r[s] = -1;
        int32_t cur = home[a];
        int nv = 0, q = 0;
        r[nv++] = cur;
        const int32_t* C = near_idx + near_ptr[a];
        const int64_t nc = near_ptr[a + 1] - near_ptr[a];
        for (int s = 0; s < 3; ++s) {
            const int32_t i = anchors[a * 3 + s];
            if (i < 0) continue;                                     /* NaN anchor, :147 */
            for (int j = 0; j < k; ++j, ++q) {
                if (!(orc_keyed_uniform(seed, 0, 5, (uint32_t)a, (uint32_t)q) >= 0.5)) continue;     /* :150 */
                const int32_t* A = in_idx + in_ptr[i];
                const int32_t* B = out_idx + out_ptr[cur];
                const int64_t na = in_ptr[i + 1] - in_ptr[i], nb = out_ptr[cur + 1] - out_ptr[cur];
                const int64_t len = orc_union_pick(A, na, B, nb, C, nc, -1, 0);
                const double u = orc_keyed_uniform(seed, 0, 6, (uint32_t)a, (uint32_t)q);
                int32_t dest = cur;
                orc_union_pick(A, na, B, nb, C, nc, (int64_t)(u * (double)len), &dest);                 /* :155 */
                cur = dest;
                r[nv++] = cur;
            }
            cur = i;                                                 /* :160-161 */
            r[nv++] = cur;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------------
 * Network construction.  TEST INFRASTRUCTURE like everything in this file.
 *
 * orc_similarity_edges -- the agent-agent edges of communities.create_network (communities.py:22-40, 65-70) without the
 * N x N matrices: for every ordered pair (i, j), i != j,
 *     p = ((((w_a * fa) * w_i) * fi) * w_d) * fd          (communities.py:38, evaluated left to right)
 *     fa = 1 - |age_j - age_i| / max|age diff|            (:29-30)
 *     fi = 1 - |inc_j - inc_i| / max|income diff|         (:24-27)
 *     fd = 1 - sqrt(dx^2 + dy^2) / max home distance      (:35-36, scipy.spatial.distance_matrix, p = 2)
 *     p  = 1 if both agents are of one household          (:39)
 * and an edge wherever p > min_prob (:65-70), rows ascending, columns ascending inside a row (np.where order).  The caller
 * forms the weight -log(p) (the reference's np.log).  Two passes: count[n] (fill = 0), then fill with rowptr = scan(count).
 * max_d2 = max over pairs of dx^2 + dy^2 (its sqrt is the largest distance: sqrt is monotone) is computed here too.
 * ------------------------------------------------------------------------------------------------------------------ */
#include <math.h>

static double pair_prob(const double* age, const double* inc, const double* x, const double* y, const int64_t* hh,
                        int64_t i, int64_t j, double max_age, double max_inc, double max_dist, double w_a, double w_i, double w_d)
{
    if (hh[i] == hh[j]) return 1.0;
    const double fa = 1.0 - fabs(age[j] - age[i]) / max_age;
    const double fi = 1.0 - fabs(inc[j] - inc[i]) / max_inc;
    const double dx = x[j] - x[i], dy = y[j] - y[i];
    const double fd = 1.0 - sqrt(dx * dx + dy * dy) / max_dist;
    return ((((w_a * fa) * w_i) * fi) * w_d) * fd;
}

int orc_similarity_edges(int64_t n, const double* age, const double* inc, const double* x, const double* y, const int64_t* hh,
                         double w_a, double w_i, double w_d, double min_prob, int fill, int64_t* count /* [n] */,
                         const int64_t* rowptr /* [n+1], fill only */, int32_t* col, double* prob, double* maxes /* [3] out */)
{
    double lo_a = age[0], hi_a = age[0], lo_i = inc[0], hi_i = inc[0], max_d2 = 0.0;
    for (int64_t i = 1; i < n; ++i) {
        if (age[i] < lo_a) lo_a = age[i];
        if (age[i] > hi_a) hi_a = age[i];
        if (inc[i] < lo_i) lo_i = inc[i];
        if (inc[i] > hi_i) hi_i = inc[i];
    }
#pragma omp parallel for schedule(dynamic, 64) reduction(max : max_d2)
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = i + 1; j < n; ++j) {
            const double dx = x[j] - x[i], dy = y[j] - y[i];
            const double d2 = dx * dx + dy * dy;
            if (d2 > max_d2) max_d2 = d2;
        }
    const double max_age = hi_a - lo_a, max_inc = hi_i - lo_i, max_dist = sqrt(max_d2);
    maxes[0] = max_age; maxes[1] = max_inc; maxes[2] = max_dist;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; ++i) {
        int64_t k = fill ? rowptr[i] : 0;
        for (int64_t j = 0; j < n; ++j) {
            if (j == i) continue;
            const double p = pair_prob(age, inc, x, y, hh, i, j, max_age, max_inc, max_dist, w_a, w_i, w_d);
            if (p > min_prob) {
                if (fill) { col[k] = (int32_t)j; prob[k] = p; }
                ++k;
            }
        }
        if (!fill) count[i] = k;
    }
    return 0;
}

/* orc_shortest_paths -- contagious3.py:124-130: scipy.sparse.csgraph.shortest_path(g, directed=True) restricted to the rows
 * of `sources`.  Dijkstra with a binary heap (lazy deletion); distances are the left-to-right sums d(v) = d(u) + w(u, v) along
 * the path, which makes the result the least fixed point of d(v) = min_u fl(d(u) + w(u,v)) -- unique, whatever the order
 * of relaxations, because rounded addition is monotone and fl(d + w) >= d for w >= 0.  out[k][v] = inf where unreachable. */
typedef struct { double d; int32_t v; } heap_item;

static void heap_push(heap_item* h, int64_t* n, double d, int32_t v)
{
    int64_t i = (*n)++;
    while (i > 0) {
        const int64_t p = (i - 1) / 2;
        if (h[p].d <= d) break;
        h[i] = h[p];
        i = p;
    }
    h[i].d = d; h[i].v = v;
}

static heap_item heap_pop(heap_item* h, int64_t* n)
{
    const heap_item top = h[0], last = h[--(*n)];
    int64_t i = 0;
    for (;;) {
        int64_t c = 2 * i + 1;
        if (c >= *n) break;
        if (c + 1 < *n && h[c + 1].d < h[c].d) ++c;
        if (last.d <= h[c].d) break;
        h[i] = h[c];
        i = c;
    }
    h[i] = last;
    return top;
}

int orc_shortest_paths(int64_t n, const int64_t* indptr, const int32_t* indices, const double* w, int64_t n_src,
                       const int32_t* sources, double* out /* [n_src][n] */)
{
    const int64_t E = indptr[n];
    for (int64_t e = 0; e < E; ++e)
        if (!(w[e] >= 0.0) || indices[e] < 0 || indices[e] >= n) return -1;
    int bad = 0;
#pragma omp parallel
    {
        heap_item* heap = (heap_item*)malloc(sizeof(heap_item) * (size_t)(E + n + 1));
#pragma omp for schedule(dynamic, 4)
        for (int64_t k = 0; k < n_src; ++k) {
            double* d = out + k * n;
            const int32_t s = sources[k];
            if (!heap || s < 0 || s >= n) { bad = 1; continue; }
            for (int64_t v = 0; v < n; ++v) d[v] = INFINITY;
            int64_t hn = 0;
            d[s] = 0.0;
            heap_push(heap, &hn, 0.0, s);
            while (hn > 0) {
                const heap_item it = heap_pop(heap, &hn);
                if (it.d > d[it.v]) continue;
                for (int64_t e = indptr[it.v]; e < indptr[it.v + 1]; ++e) {
                    const double nd = it.d + w[e];
                    if (nd < d[indices[e]]) { d[indices[e]] = nd; heap_push(heap, &hn, nd, indices[e]); }
                }
            }
        }
        free(heap);
    }
    return bad ? -1 : 0;
}

/* orc_building_edges -- the building-building edges of communities.create_network (communities.py:12-21, 53-54): the gravity
 * model  scores[i][j] = fs[j] / dist(i, j)^beta  with beta = 2 (dist^2 is the SQUARE OF THE ROUNDED DISTANCE, as
 * `distance_matrix(...)**beta` forms it; a zero becomes 0.00001, :14), scores[i][i] = 0, prob = scores / row sum, an edge
 * i -> j wherever prob > min_prob and i != j.  The row sum is numpy's `scores.sum(axis=1)`: PAIRWISE summation over the
 * contiguous row (blocks of <= 128 elements with eight strided accumulators, halves split at a multiple of 8) -- restated
 * here because the order of the additions decides the last bit of every probability. */
static double bld_score(const double* fs, const double* x, const double* y, int64_t i, int64_t j)
{
    if (i == j) return 0.0;
    const double dx = x[j] - x[i], dy = y[j] - y[i];
    const double d = sqrt(dx * dx + dy * dy);
    double d2 = d * d;
    if (d2 == 0.0) d2 = 0.00001;
    return fs[j] / d2;
}

static double bld_pairwise(const double* fs, const double* x, const double* y, int64_t i, int64_t j0, int64_t n)
{
    if (n < 8) {
        double res = 0.;                       /* (numpy starts from 0. here; -0.0 cannot occur: scores are >= 0) */
        for (int64_t k = 0; k < n; ++k) res += bld_score(fs, x, y, i, j0 + k);
        return res;
    }
    if (n <= 128) {
        double r[8];
        for (int q = 0; q < 8; ++q) r[q] = bld_score(fs, x, y, i, j0 + q);
        int64_t k;
        for (k = 8; k < n - (n % 8); k += 8)
            for (int q = 0; q < 8; ++q) r[q] += bld_score(fs, x, y, i, j0 + k + q);
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; k < n; ++k) res += bld_score(fs, x, y, i, j0 + k);
        return res;
    }
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    return bld_pairwise(fs, x, y, i, j0, n2) + bld_pairwise(fs, x, y, i, j0 + n2, n - n2);
}

int orc_building_edges(int64_t B, const double* fs, const double* x, const double* y, double min_prob, int fill,
                       int64_t* count /* [B] */, const int64_t* rowptr, int32_t* col, double* prob, double* row_sum /* [B] out, nullable */)
{
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < B; ++i) {
        const double s = bld_pairwise(fs, x, y, i, 0, B);
        if (row_sum) row_sum[i] = s;
        int64_t k = fill ? rowptr[i] : 0;
        for (int64_t j = 0; j < B; ++j) {
            const double p = bld_score(fs, x, y, i, j) / s;
            if (p > min_prob && j != i) {
                if (fill) { col[k] = (int32_t)j; prob[k] = p; }
                ++k;
            }
        }
        if (!fill) count[i] = k;
    }
    return 0;
}
