// Developer micro-benchmark (not part of the product): times kernels on synthetic arrays to separate launch floor,
// latency chains and throughput.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo tools/kbench.cu -o tools/kbench
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../dysturbd_b200/csrc/dys_day.cuh"
using namespace dys;

__global__ void k_empty(const DevCtx c, const int day) {}
__global__ void __launch_bounds__(256) k_stream(const DevCtx c, const int day)
{
    const int64_t a0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    if (a0 >= c.n_pad) return;
    const uchar4 sv = *reinterpret_cast<const uchar4*>(c.status + a0);
    uint32_t w = sv.x + sv.y + sv.z + sv.w + c.rec[a0].t_inf;
    *reinterpret_cast<uint32_t*>(c.ib_wr[0] + a0) = w;
}
__global__ void __launch_bounds__(256) k_agent_nofold(const DevCtx c, const int day)
{
    __shared__ BlockScratch sm;
    unsigned int none[4] = {0, 0, 0, 0};
    agent_body<1>(c, day, sm, nullptr, none, (int)gridDim.x, (int)blockIdx.x);
}
__global__ void __launch_bounds__(256) k_agent_noreduce(const DevCtx c, const int day)
{
    __shared__ BlockScratch sm;
    unsigned int none[4] = {0, 0, 0, 0};
    agent_body<2>(c, day, sm, nullptr, none, (int)gridDim.x, (int)blockIdx.x);
}
template <typename T> T* dalloc(size_t n, int fill = 0) { T* p; cudaMalloc(&p, n * sizeof(T)); cudaMemset(p, fill, n * sizeof(T)); return p; }

int main(int argc, char** argv)
{
    const int64_t N = argc > 1 ? atoll(argv[1]) : 1002496;
    const double frac_inf = argc > 2 ? atof(argv[2]) : 0.02;
    DevCtx c{};
    c.n_loc = c.n_pad = c.n_glob = N; c.rt_n = N; c.H = N / 3; c.B = N / 20; c.S = N / 1200; c.V = 10; c.B_out = c.B; c.S_out = c.S;
    c.recover = 21; c.hospital_recover = 28; c.diagnosis = 7; c.quarantine = 7; c.adm_min = 4; c.adm_max = 14; c.death_min = 3;
    c.norm_factor = 4.0; c.compliance = 1.0; c.seed = 1;
    std::vector<uint8_t> st(N, ST_S); std::vector<int16_t> ti(N, 0); std::vector<int32_t> hh(N), home(N), sa(N), src(N, -1);
    srand(1);
    for (int64_t a = 0; a < N; ++a) {
        hh[a] = (int)(a / 3); home[a] = (int)(a / 30 % c.B); sa[a] = (int)(a / 1200 % c.S);
        if (rand() / (double)RAND_MAX < frac_inf) { st[a] = ST_I; ti[a] = 18; src[a] = 5; }
    }
    c.status = dalloc<uint8_t>(N); cudaMemcpy(c.status, st.data(), N, cudaMemcpyHostToDevice);
    {
        std::vector<AgentRec> rc(N);
        for (int64_t a = 0; a < N; ++a) rc[a] = AgentRec{ti[a], 0, 0, 3, sa[a], home[a], hh[a], (int32_t)(a / 3 * 3), 0.5};
        c.rec = dalloc<AgentRec>(N); cudaMemcpy(c.rec, rc.data(), sizeof(AgentRec) * N, cudaMemcpyHostToDevice);
    }
    c.src = dalloc<int32_t>(N); cudaMemcpy(c.src, src.data(), 4 * N, cudaMemcpyHostToDevice);
    int32_t* d; d = dalloc<int32_t>(N); cudaMemcpy(d, hh.data(), 4 * N, cudaMemcpyHostToDevice); c.hh = d;
    c.risk = dalloc<double>(N); c.p_adm = dalloc<double>(N); c.p_death = dalloc<double>(N);
    c.pdf = dalloc<double>(64);
    DayBuffers buf{};
    c.n_peers = 1;
    c.flag_wait = dalloc<int32_t>(8); c.flag_signal[0] = dalloc<int32_t>(8);
    c.open = dalloc<uint8_t>(c.B, 1); c.n_closed = 0;
    for (int k = 0; k < 2; ++k) {
        buf.att[k] = dalloc<uint8_t>(N + 16); buf.mb_any[k] = dalloc<int32_t>(N, 0xFF); buf.mb_iso[k] = dalloc<int32_t>(N, 0xFF);
        buf.ib[k] = dalloc<uint8_t>(N + 2048);
    }
    for (int k = 0; k < Q_RING; ++k) buf.qflag[k] = dalloc<uint8_t>(c.H + 16);
    buf.df = dalloc<int32_t>(DF_RING * DF_LEN);
    buf.part = dalloc<unsigned long long>(PART_RING * 64 * 64);
    int32_t* df = buf.df;
    uint8_t* ibbuf = buf.ib[0];
    c.stats = dalloc<int64_t>(64); c.ticket = dalloc<unsigned int>(1);
    {
        std::vector<int64_t> hp(c.H + 2); std::vector<int32_t> hm(N);
        for (int64_t h = 0; h <= c.H + 1; ++h) hp[h] = h * 3 < N ? h * 3 : N;
        for (int64_t a = 0; a < N; ++a) hm[a] = (int)a;
        int64_t* dp = dalloc<int64_t>(c.H + 2); cudaMemcpy(dp, hp.data(), 8 * (c.H + 2), cudaMemcpyHostToDevice); c.hh_ptr = dp;
        int32_t* dm = dalloc<int32_t>(N); cudaMemcpy(dm, hm.data(), 4 * N, cudaMemcpyHostToDevice); c.hh_mem = dm;
    }
    bind_day(c, buf, 100);
    c.sa_hist = dalloc<int32_t>(c.S * 21); c.bcount = dalloc<int32_t>(c.B);
    c.ev = dalloc<int2>(N);
    {   // a synthetic transposed network: 17 entries per column, rows within +-300 of the column (same community)
        const int D = 17;
        std::vector<int64_t> tp(N + 1);
        std::vector<uint32_t> tc((size_t)N * D + 16);
        std::vector<double> tv((size_t)N * D + 2, 0.97);
        for (int64_t i = 0; i <= N; ++i) tp[i] = i * D;
        for (int64_t i = 0; i < N; ++i)
            for (int k = 0; k < D; ++k) {
                int64_t u = i + (rand() % 601) - 300;
                if (u < 0) u = 0; if (u >= N) u = N - 1;
                tc[(size_t)i * D + k] = (uint32_t)u | ((rand() % 5 == 0 ? 1u : 0u) << 28);
            }
        int64_t* dp = dalloc<int64_t>(N + 1); cudaMemcpy(dp, tp.data(), 8 * (N + 1), cudaMemcpyHostToDevice); c.tptr = dp;
        std::vector<TEdge> te((size_t)N * D + 4);
        for (size_t e = 0; e < (size_t)N * D; ++e) { te[e].colx = tc[e]; te[e].pad = 0; te[e].val = tv[e] * 0.1; }
        TEdge* de = dalloc<TEdge>(te.size()); cudaMemcpy(de, te.data(), sizeof(TEdge) * te.size(), cudaMemcpyHostToDevice); c.tedge = de;
        std::vector<double> rk(N, 0.1); cudaMemcpy((void*)c.risk, rk.data(), 8 * N, cudaMemcpyHostToDevice);
        std::vector<double> pdf(64, 0.1); cudaMemcpy((void*)c.pdf, pdf.data(), 8 * 64, cudaMemcpyHostToDevice);
        c.norm_factor = 1e-6;       // Bernoulli successes are rare: the mailbox stays (almost) untouched between launches
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](const char* name, auto launch) {
        for (int i = 0; i < 5; ++i) launch(i);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        const int R = 200;
        for (int i = 0; i < R; ++i) launch(100);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-28s %8.2f us/launch   (%s)\n", name, ms * 1000 / R, cudaGetErrorString(cudaGetLastError()));
    };
    const unsigned g4 = (unsigned)(N / 1024);
    int anynd = 1;
    timeit("k_empty", [&](int d) { k_empty<<<g4, 256>>>(c, d); });
    timeit("k_stream (5 B/agent)", [&](int d) { k_stream<<<g4, 256>>>(c, d); });
    timeit("k_agent any_nd=0", [&](int d) { k_agent<<<g4, 256>>>(c, d); });
    cudaMemcpy(df, &anynd, 4, cudaMemcpyHostToDevice);
    timeit("k_agent any_nd=1", [&](int d) { k_agent<<<g4, 256>>>(c, d); });
    timeit("k_agent no fold/ticket", [&](int d) { k_agent_nofold<<<g4, 256>>>(c, d); });
    timeit("k_agent no reduction", [&](int d) { k_agent_noreduce<<<g4, 256>>>(c, d); });
    timeit("k_snapshot", [&](int d) { k_snapshot<<<g4, 256>>>(c, d); });
    int per_sm = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_push, 256, 0);
    {   // snapshot bytes for the infectious agents
        std::vector<uint8_t> ibh(N + 2048, 0);
        for (int64_t a = 0; a < N; ++a) if (st[a] == ST_I) ibh[a] = 0x80 | 3;
        cudaMemcpy(ibbuf, ibh.data(), N, cudaMemcpyHostToDevice);
    }
    timeit("k_push", [&](int d) { k_push<<<148 * per_sm, 256>>>(c, d); });
    printf("push blocks/SM %d\n", per_sm);
    {   // k_days: 7 days in one cooperative launch, with phase timestamps of block 0
        DayPlan plan{};
        plan.first_day = 100; plan.n_days = 7;
        plan.off_sa = 512; plan.off_bc = 512 + ((c.S * 21 * 4 + 15) & ~15ll); plan.off_ev = plan.off_bc + ((c.B * 4 + 15) & ~15ll);
        plan.want_sa = plan.want_bc = plan.want_ev = 1;
        const size_t blk = plan.off_ev + 8 * (size_t)N;
        for (int k = 0; k < 7; ++k) { plan.out[k] = dalloc<unsigned char>(blk); plan.open[k] = c.open; plan.n_closed[k] = 0; }
        cudaMemcpy(buf.ib[1], ibbuf, N, cudaMemcpyDeviceToDevice);
        plan.buf = buf;
        plan.trace = dalloc<unsigned long long>((size_t)7 * 148 * 8 * TRACE_SLOTS);
        unsigned int* bar = dalloc<unsigned int>(4);
        int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_days, 256, 0);
        void* args[] = {(void*)&c, (void*)&plan, (void*)&bar};
        for (int rep = 0; rep < 3; ++rep) {
            cudaMemset(bar, 0, 16);
            for (int q = 0; q < DF_RING; ++q) cudaMemcpy(df + q * DF_LEN, &anynd, 4, cudaMemcpyHostToDevice);
            cudaEventRecord(e0);
            cudaError_t er = cudaLaunchCooperativeKernel((const void*)k_days, dim3(148 * occ), dim3(256), args, 0, 0);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("k_days 7 days: %.2f us/day (%s, %d CTAs/SM)\n", ms * 1000 / 7, cudaGetErrorString(er), occ);
        }
        const int G = 148 * occ;
        std::vector<unsigned long long> tr((size_t)7 * G * TRACE_SLOTS);
        cudaMemcpy(tr.data(), plan.trace, tr.size() * 8, cudaMemcpyDeviceToHost);
        for (int k = 0; k < 7; ++k) {
            // per phase: mean over the working blocks, and the latest block relative to the earliest day start
            double sw = 0, se = 0, pu = 0, late = 0; unsigned long long t_first = ~0ull, t_last = 0;
            for (int b = 0; b < G - 1; ++b) {
                const unsigned long long* t = &tr[((size_t)k * G + b) * TRACE_SLOTS];
                sw += (t[5] - t[2]) / 1e3; se += (t[6] - t[5]) / 1e3; pu += (t[7] - t[6]) / 1e3;
                if (t[2] < t_first) t_first = t[2];
                if (t[3] > t_last) t_last = t[3];
            }
            (void)late;
            const unsigned long long* t0 = &tr[((size_t)k * G) * TRACE_SLOTS];
            printf("  day %d: sweep %.2f  settle %.2f  push %.2f us (means)   first start -> last block done %.2f   block 0 barrier wait %.2f\n",
                   k, sw / (G - 1), se / (G - 1), pu / (G - 1), (t_last - t_first) / 1e3, (t0[4] - t0[3]) / 1e3);
        }
    }
    return 0;
}
