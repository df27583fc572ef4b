#!/usr/bin/env python
"""bench.py -- agent-days/s of the per-day epidemic step (BASELINE.json metric) on N B200s of one node.

  python bench.py --gpus 1 --steps K --warmup W            N=1: the synthetic ~1 M-agent city (BASELINE.json configs[1])
  torchrun ... bench.py --gpus N --steps K --warmup W      N>1: ONE coupled city of N x ~1 M agents, sharded by community,
                                                           the per-day exchange fused into the day kernel (weak scaling)
  ... --scaling strong --agents-total 10002432             configs[2]: a fixed city split over the N ranks
  ... --communities-per-gpu 1110                           configs[3]: 12.5 M agents per GPU (100 M on 8)
  ... --replicas R                                         configs[4]: R Monte-Carlo replicas of the city per GPU, batched
                                                           into one launch, no communication
  ... --scenario GRADUAL,ALL                               a lockdown scenario (flags decided on the host every day)
  python bench.py --impl reference ...                     the CPU path on the host cores, same config and metric

A step is one simulated day (one pass of contagious3.py:176-366) over the whole population.  Prints ONE JSON line.
`value`   : days W..W+K-1 back to back with the population resident in HBM, timed with CUDA events recorded on the
            library's stream right around the launches.  The K days are replayed (state reset and W warm-up days untimed)
            until ~0.5 s of timed work or 40 replays; value is the MEDIAN replay, the spread is reported.
`e2e`     : the same days through the documented host API -- dysturbd_b200.host.EpidemicModel(compact=True).run(): every
            window the day's building flags go host -> device from pinned memory, every day's output block (stats,
            per-area histograms, per-building counts, infection events) comes device -> host and the reference's Stats /
            SAs / Buildings / IO_mat entries are built from it.  Wall clock, median over replays.
`roofline`: k_days (the fused day kernel): SURVEY.md 8(d)'s algorithmic bytes per day evaluated with the days' own
            counters, over the mean duration of a day in the resident pass, against MEASURED_PEAKS.json; next to it the
            bytes THIS design actually has to move and what ncu measured.
`cpu_baseline`: the C oracle (OpenMP, all host cores) on the same population, a bounded sample of days, rank 0, N=1.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WINDOW = 7


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=365)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--communities-per-gpu", type=int, default=89, help="89 x 11 264 = 1 002 496 agents per GPU")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--agents-total", type=int, default=0, help="strong scaling: total agents (a multiple of 11 264 x N)")
    ap.add_argument("--replicas", type=int, default=1, help="Monte-Carlo replicas of the city per GPU, one launch for all")
    ap.add_argument("--scenario", default="noLockdown", help="comma separated scenario keywords, e.g. GRADUAL,ALL")
    ap.add_argument("--norm-factor", type=float, default=6.0,
                    help="with the degree-16.7 synthetic network 6.0 gives a reference-like epidemic: peak near day 70, "
                         "38 %% of the population infected, extinct near day 280")
    ap.add_argument("--compliance", type=float, default=1.0)
    ap.add_argument("--seeds-per-22493", type=int, default=20)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cpu-baseline-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the batched-replica and lockdown side measurements of the N=1 line")
    ap.add_argument("--max-replays", type=int, default=40)
    ap.add_argument("--verify-sharded", action="store_true", help="N>2: also check the sharded run against a single-GPU run (always done at N=2)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N>1: the exchange fused into k_days over peer memory (peer) or a per-day NCCL all-gather")
    return ap.parse_args()


def communities(args, world):
    from dysturbd_b200.synth import COMMUNITY
    if args.scaling == "strong":
        total = args.agents_total or COMMUNITY * 888
        n_comm = total // COMMUNITY
        if n_comm % world or n_comm * COMMUNITY != total:
            raise SystemExit("--agents-total must be a multiple of 11264 x N")
        return n_comm // world
    return args.communities_per_gpu


def make_population(args, rank, world):
    from dysturbd_b200.synth import COMMUNITY, synth_population
    cpg = communities(args, world)
    n_total = COMMUNITY * cpg * world
    comm = None if world == 1 else (rank * cpg, (rank + 1) * cpg)
    pop = synth_population(n_total, seed=args.seed, communities=comm, seeds_per_22493=args.seeds_per_22493)
    return pop, n_total


def make_params(args):
    from dysturbd_b200.population import EpiParams
    return EpiParams(norm_factor=args.norm_factor, compliance=args.compliance, seed=20201019)


def workload_string(args, world, n_total):
    """identical in both arms, so that the driver sees the same config"""
    cpg = communities(args, world)
    return ("synthetic city, %d agents (%d GPUs x %d communities of 11264; ~300k households, ~50k buildings per million), "
            "communities social network (degree ~16.7), norm_factor %g, scenario %s, %d replica(s) per GPU, days %d..%d of "
            "one epidemic" % (n_total, world, cpg, args.norm_factor, args.scenario, args.replicas, args.warmup,
                              args.warmup + args.steps - 1))


class ClockSampler:
    """SM clocks and throttle reasons while the GPU is under the benchmark's load, polled through NVML from a thread.
    Nothing here runs between two CUDA-event records: the thread only samples while `active`, and `active` is only
    set around regions that are timed by wall clock or not timed at all."""

    def __init__(self, gpu_index):
        self.rows, self.gpu, self.active, self._stop = [], gpu_index, False, False
        self.nv = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
            return
        threading.Thread(target=self._poll, daemon=True).start()

    def _poll(self):
        nv = self.nv
        while not self._stop:
            if self.active:
                try:
                    sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                    try:
                        rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    self.rows.append((sm, rs))
                except Exception:
                    pass
            time.sleep(0.004)

    def stop(self):
        self._stop = True

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4,
                 "hw_power_brake": 0x80}
        seen = set()
        for _, rs in self.rows:
            for n, bit in names.items():
                if rs & bit:
                    seen.add(n)
        return {"sm_mhz": float(np.median([r[0] for r in self.rows])), "sm_max_mhz": float(self.max_sm), "reasons": sorted(seen),
                "samples": len(self.rows)}


def design_bytes(N, st, prev, window=WINDOW):
    """bytes one simulated day must move in THIS design (DESIGN.md 'bytes per day'), from the day's own counters: each
    needed datum once, at its stored width.  `st` = today's stats block, `prev` = yesterday's (its end-of-day status
    counts are today's day-start counts)."""
    from dysturbd_b200.engine import STAT
    g = lambda k: int(st[STAT[k]])
    S0, I0, QH0, QI0, D0, H0, R0, X0 = (int(x) for x in prev[20:28])
    rows = g("inf0")
    push = rows * 16 + g("edges_scanned") * 16 + g("hits") * (4 + 4 + 1) + N // window
    listed = I0 + QI0 + D0 + H0 + X0 + QH0 + g("hits") + g("daily_quar")
    agent = N * (1 + 1) + listed * 16
    agent += g("changed") * (1 + 2) + g("n_events") * (8 + 4) + g("daily_quar") * 2
    agent += int(st[32:53].sum()) * 4 + (I0 + QI0 + D0) * 4      # per-area histogram bin, per-building count
    agent += 2 * N // window                                    # tomorrow's snapshot bytes: last day of a window only
    return push + agent


def survey_8d_bytes(N, st, prev, mean_degree):
    """SURVEY.md section 8(d)'s algorithmic bytes per day (compact SoA widths, each needed datum once), evaluated with
    the day's counters."""
    from dysturbd_b200.engine import STAT
    g = lambda k: int(st[STAT[k]])
    S0, I0, QH0, QI0, D0, H0, R0, X0 = (int(x) for x in prev[20:28])
    n_live = S0 + I0 + QH0 + QI0 + D0
    e_scanned = int(g("inf0") * mean_degree)                     # every interaction_prob non-zero of an infectious agent
    u_gated = min(S0 + QH0, g("exposed"))
    b = N * 2 + N * 6 + n_live * 40 + (I0 + QI0 + D0) * 8 + H0 * 8 + u_gated * 16 + e_scanned * 12
    b += N * 4 if (g("new_diag") or g("daily_quar")) else 0
    b += N * 6 + g("changed") * 11
    return b


def run_ours(args):
    import torch
    import torch.distributed as dist
    from dysturbd_b200 import engine
    from dysturbd_b200.engine import STAT
    from dysturbd_b200.host import EpidemicModel
    from dysturbd_b200.sharded import ShardedEngine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus must equal WORLD_SIZE")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local)
    gloo = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # the host loop's per-window sum of 64-integer stats blocks goes over gloo (host memory): a collective KERNEL would
        # have to find room on a GPU whose SMs are all held by the persistent day kernel
        gloo = dist.new_group(backend="gloo")
    pop, n_total = make_population(args, rank, world)
    prm = make_params(args)
    W, K, R = args.warmup, args.steps, args.replicas
    scenario = args.scenario.split(",")
    lockdown = any(k in scenario for k in ("ALL", "EDU", "REL"))
    all_out = engine.WANT_SA_HIST | engine.WANT_BUILDING_COUNTS | engine.WANT_EVENTS
    replicas_mode = R > 1                # (at N>1 every rank runs its own batch of whole cities, no exchange)
    if replicas_mode:
        pop, _ = make_population(args, 0, 1)          # every rank holds R replicas of the SAME whole city
        n_city = pop.N
        n_total = n_city * R * world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    t0 = time.perf_counter()
    if replicas_mode:
        eng = engine.Engine(pop, prm, device=local, flags=0, replicas=R)
        # the parameter sweep of configs[4]: a grid over infection rate x quarantine compliance, one seed per replica
        g = int(np.ceil(np.sqrt(R * world)))
        idx = np.arange(R) + rank * R
        norms = np.linspace(0.5, 1.5, g)[idx % g] * args.norm_factor
        comps = np.linspace(0.25, 1.0, g)[(idx // g) % g]
        eng.set_replica_params(norms, comps, 20201019 + idx)
        raw = eng
    elif world == 1:
        # the day's building flags are this path's only per-day HOST input (contagious3.py:384 re-assigns build[:,10] every
        # day): in the end-to-end pass every day's vector travels, changed or not (the library would skip unchanged ones)
        os.environ["DYS_UPLOAD_ALWAYS"] = "1"
        eng = raw = engine.Engine(pop, prm, device=local, flags=all_out)
        del os.environ["DYS_UPLOAD_ALWAYS"]
    else:
        eng = ShardedEngine(pop, prm, device=local, mode=args.exchange, flags=all_out)
        raw = eng.sim
    eng_mode = getattr(eng, "mode", "single")
    raw.sync()
    load_s = time.perf_counter() - t0
    stream = torch.cuda.ExternalStream(raw.cuda_stream(), device=local)
    init = pop.copy_state()

    def reset():
        eng.set_state(init["status"], init["t_inf"], init["t_q"], init["t_adm"], init["src"])
        barrier()                        # (peer mode: nobody steps before every rank has reset its flags)

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()

    # ------------------------------------------------------------------------------------------------------------
    # pass A: end to end through the documented host API (also records the flags and counters of every day)
    # ------------------------------------------------------------------------------------------------------------
    def reduce_stats(block):
        t = torch.from_numpy(np.array(block, dtype=np.int64))
        dist.all_reduce(t, group=gloo)
        return t.numpy()

    def e2e_once(sc, collect):
        """one replay: W warm-up days and K timed days through EpidemicModel.run()"""
        reset()
        m = EpidemicModel(None, None, None, None, scenario=sc, params=prm, population=pop, compact=True,
                          backend=lambda p_, q_: eng, reduce_stats=reduce_stats if world > 1 else None,
                          upload_flags=(world == 1), keep_raw=collect, choice_seed=7)
        b0 = raw.h2d_bytes()
        m.run(max_days=W, finalize=False)
        barrier()
        b1 = raw.h2d_bytes()
        t = time.perf_counter()
        m.run(max_days=W + K, finalize=False)
        dt = time.perf_counter() - t
        done = m.day - W                  # (an epidemic that dies out before W + K ends the run, as in the reference)
        out = {"seconds": max_over_ranks(dt), "days": done, "h2d": raw.h2d_bytes() - b1}
        if collect:
            out["flags"] = m.flags_log if lockdown else None
            out["raw"] = m.raw_stats
            out["model"] = m
        return out

    def replicas_e2e_once():
        """replicas: the host loop of a sweep -- windows of days for all replicas, every replica's stats block read each day"""
        reset()
        d = 0
        pinned = torch.ones((WINDOW, pop.B), dtype=torch.uint8).pin_memory().numpy()
        tot = 0
        t = None
        while d < W + K:
            if d >= W and t is None:
                raw.sync()
                barrier()
                t = time.perf_counter()
            n = min(WINDOW, (W if d < W else W + K) - d)
            raw.step_many(d, n, pinned[:n])
            for r in range(R):
                raw.select_replica(r)
                for dd in range(d, d + n):
                    tot += int(raw.outputs(dd)["stats"][0])
            d += n
        raw.select_replica(-1)
        raw.sync()
        return {"seconds": max_over_ranks(time.perf_counter() - t), "days": K, "checksum": tot}

    e2e_runs = []
    first = None
    clocks.active = True
    n_e2e = 0
    while True:
        r = replicas_e2e_once() if replicas_mode else e2e_once(scenario, collect=first is None)
        if first is None:
            first = r
        e2e_runs.append(r["seconds"] / max(1, r["days"]))
        n_e2e += 1
        if n_e2e >= args.max_replays or (n_e2e >= 3 and sum(e2e_runs) * K > 0.5) or n_e2e * r["seconds"] > 20:
            break
    clocks.active = False
    s_e2e_day = float(np.median(e2e_runs))
    days_run = first["days"]
    m0 = first.get("model")
    flags_rec = first.get("flags")
    counters = sorted(first["raw"].items()) if first.get("raw") else None

    # ------------------------------------------------------------------------------------------------------------
    # pass B: resident throughput (`value`): CUDA events on the library's stream, right around the launches
    # ------------------------------------------------------------------------------------------------------------
    last_flags = [None]

    def flags_of(d, n):
        """the recorded flag vectors of days d .. d+n-1, or None when none of them differs from the vector in force (the
        library keeps the last vector of a window as the standing one): unchanged flags are not uploaded again"""
        if flags_rec is None:
            return None
        fl = [flags_rec[min(x, max(flags_rec))] for x in range(d, d + n)]
        if d > 0 and last_flags[0] is not None and all(f is last_flags[0] or np.array_equal(f, last_flags[0]) for f in fl):
            return None
        last_flags[0] = fl[-1]
        return np.stack(fl)

    def run_resident(first_day, last_day, win):
        d = first_day
        while d < last_day:
            n = min(win, last_day - d)
            eng.step_many(d, n, flags_of(d, n))
            d += n

    Kv = days_run if days_run > 0 else K          # (the days the epidemic actually lasted, if it died out early)
    win = WINDOW
    res_ms = []
    with torch.cuda.stream(stream):
        while True:
            reset()
            run_resident(0, W, win)
            raw.sync()
            barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            run_resident(W, W + Kv, win)
            ev1.record(stream)           # (recorded straight after the last launch: no host call in between)
            raw.sync()
            barrier()
            res_ms.append(max_over_ranks(ev0.elapsed_time(ev1)))
            if len(res_ms) >= args.max_replays or (len(res_ms) >= 3 and sum(res_ms) > 500.0) or sum(res_ms) > 20000.0:
                break
    ms_value = float(np.median(res_ms))

    # clocks under load: an untimed stretch of the same stepping with the NVML poller on
    clocks.active = True
    t_end = time.perf_counter() + 0.4
    with torch.cuda.stream(stream):
        while time.perf_counter() < t_end:
            reset()
            run_resident(0, W + Kv, win)
            raw.sync()
    clocks.active = False

    # ------------------------------------------------------------------------------------------------------------
    # sharded == unsharded (outside every timed region): rank 0 runs the WHOLE city on its own GPU
    # ------------------------------------------------------------------------------------------------------------
    verified = None
    if world > 1 and not replicas_mode and (world == 2 or args.verify_sharded):
        glob = np.stack([c[1] for c in counters]) if counters else np.zeros((0, 64), np.int64)
        if rank == 0:
            from dysturbd_b200.synth import synth_population
            whole = synth_population(n_total, seed=args.seed, seeds_per_22493=args.seeds_per_22493)
            single = engine.Engine(whole, prm, device=local, flags=0)
            ok = True
            for d in range(W + Kv):
                single.step_many(d, 1, flags_of(d, 1))
                hit = [c for c in counters if c[0] == d]
                if hit:
                    s1, sN = single.stats(d), hit[0][1]
                    keep = [i for i in range(14)] + list(range(20, 28)) + list(range(32, 53))
                    ok = ok and bool(np.array_equal(s1[keep], sN[keep]))
            single.close()
            verified = {"days_compared": len(counters), "equal": ok}
            assert ok, "the sharded run diverged from the single-GPU run of the same city"
        barrier()

    # ---- byte formulas from the days' counters ----
    survey_b = design_b = 0
    n_counted = 0
    if counters:
        by_day = dict(counters)
        for d in range(W, W + Kv):
            if d in by_day and d - 1 in by_day:
                survey_b += survey_8d_bytes(n_total, by_day[d], by_day[d - 1], pop.nnz / pop.N)
                design_b += design_bytes(n_total, by_day[d], by_day[d - 1])
                n_counted += 1
    st_last = counters[-1][1] if counters else None

    # ---- N=1 extras: replicas batched in one launch, and the same city under a lockdown scenario ----
    extras = {}
    if world == 1 and not replicas_mode and not args.no_extras:
        eng.close()
        extras = run_extras(args, pop, prm, local, stream_of=lambda e: torch.cuda.ExternalStream(e.cuda_stream(), device=local))
    else:
        eng.close()
    clocks.stop()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    k_ms = ms_value / max(1, Kv)
    per_gpu = lambda b: b / max(1, n_counted) / world
    roof = None
    if n_counted:
        achieved = per_gpu(survey_b) / (k_ms * 1e-3) / 1e9
        d_ach = per_gpu(design_b) / (k_ms * 1e-3) / 1e9
        roof = {"bound": "hbm",
                "kernel": "k_days (per simulated day: sweep + settle of the block's pieces, push of the block's own frontier for "
                          "the next day, ONE grid barrier; one cooperative launch per %d days)" % win,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                "traffic": None,
                "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum is not measurable inside this process; the ncu "
                                "--set full capture of the same command is summarised under profiles/ (r02*_summary.md)",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 (B200_PROFILING.md)",
                "algorithmic_bytes_per_launch": per_gpu(survey_b) * win, "ms_per_launch": k_ms * win,
                "bytes_definition": "SURVEY.md 8(d): per-day bytes of a compact-SoA pull (status, day columns, routine rows of "
                                    "live agents, gated interaction_prob rows at 12 B per non-zero, state writes), evaluated "
                                    "with each day's own counters, per GPU",
                "this_design": {"bytes_per_day": per_gpu(design_b), "achieved": d_ach, "frac": d_ach / peak if peak else None,
                                "what": "the bytes THIS implementation has to move per day (static pair classes instead of "
                                        "routine rows, only the transposed rows of infectious agents, 2 bytes per idle agent): "
                                        "DESIGN.md section 4; ncu's dram bytes agree with this figure, not with 8(d)'s"}}
    value = n_total * Kv / (ms_value * 1e-3)
    d2h_day = 0
    if not replicas_mode:
        o = m0.sim.sim if world > 1 else m0.sim
        d2h_day = 64 * 8 + o.S_out * 21 * 4 + o.B_out * 4 + max(1024, ((pop.N + 1023) // 1024 * 1024) // 64) * 8
    else:
        d2h_day = 64 * 8 * R
    line = {
        "metric": "agent-days/sec", "value": value, "unit": "agent-days/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": k_ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "u8/i16 state, f64 probabilities", "data": "synthetic",
        "config": {
            "workload": workload_string(args, world, n_total if not replicas_mode else pop.N),
            "agents_total": n_total, "agents_per_gpu": pop.N * (R if replicas_mode else 1), "edges_per_gpu": pop.nnz,
            "norm_factor": args.norm_factor, "seeds_per_22493": args.seeds_per_22493, "scenario": args.scenario,
            "replicas_per_gpu": R, "days_run": Kv,
            "parallelism": ("1 GPU" if world == 1 else
                            "%d GPUs, %d independent replicas each, no communication" % (world, R) if replicas_mode else
                            "sharded by community x%d, exchange=%s (peer: snapshot bytes stored into the peer GPUs by the day "
                            "kernel over NVLink; nccl: per-day all-gather)" % (world, eng_mode)),
            "l2": "no flush: a day's working set (state + network, %.0f MB per GPU) is larger than the 126 MB L2 and every day "
                  "rewrites the state" % (eng_bytes(pop) * (R if replicas_mode else 1) / 1e6),
            "timing": "value = median of %d replays of the %d days (CUDA events on the library stream); spread min %.4f / "
                      "max %.4f ms per day" % (len(res_ms), Kv, min(res_ms) / max(1, Kv), max(res_ms) / max(1, Kv)),
            "ever_infected_at_end": int(st_last[STAT["ever_infected"]]) if st_last is not None else None,
            "load_seconds": load_s,
            "sharded_equals_single_gpu": verified,
        },
        "e2e": {"value": n_total * 1.0 / s_e2e_day, "unit": "agent-days/s",
                "h2d_bytes_per_step": float(pop.B) if replicas_mode else first["h2d"] / max(1, days_run),
                "d2h_bytes_per_step": float(d2h_day), "ms_per_step": s_e2e_day * 1e3, "replays": len(e2e_runs),
                "api": ("dysturbd_b200.host.EpidemicModel(compact=True).run(): flags host->device every window from pinned "
                        "memory, every day's output block device->host, Stats/SAs/Buildings/IO_mat built on the host"
                        if not replicas_mode else
                        "Engine.step_many + per-replica Engine.outputs(): flags host->device every window, every replica's "
                        "stats block device->host every day")},
        "gpu_launches": -(-Kv // win),  # one cooperative k_days launch per window (k_snapshot only after a state load)
        "roofline": roof,
        "clocks": clocks.summary(),
    }
    line.update(extras)
    if world == 1 and not args.no_cpu_baseline and not replicas_mode:
        line["cpu_baseline"] = cpu_baseline(pop, prm, args.cpu_baseline_seconds, W, K, scenario=scenario)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_extras(args, pop, prm, local, stream_of):
    """Side measurements of the default N=1 line (same city, same process): (1) R replicas batched into one launch
    (BASELINE.json configs[4]) -- the independent work that fills the GPU; (2) the documented host API under the
    reference's own default scenario ['GRADUAL', 'ALL'] (parameters.py:53), where the flags change and really travel."""
    import torch
    from dysturbd_b200 import engine
    from dysturbd_b200.host import EpidemicModel
    out = {}
    init = pop.copy_state()
    W, K = args.warmup, min(args.steps, 56)
    try:
        Rr = 16
        e = engine.Engine(pop, prm, device=local, flags=0, replicas=Rr)
        e.set_replica_params(None, None, 20201019 + np.arange(Rr))
        st = stream_of(e)
        ms = []
        with torch.cuda.stream(st):
            for _ in range(5):
                e.set_state(init["status"], init["t_inf"], init["t_q"], init["t_adm"], init["src"])
                d = 0
                while d < W:
                    n = min(WINDOW, W - d); e.step_many(d, n); d += n
                e.sync()
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record(st)
                while d < W + K:
                    n = min(WINDOW, W + K - d); e.step_many(d, n); d += n
                ev1.record(st)
                e.sync()
                ms.append(ev0.elapsed_time(ev1))
        m = float(np.median(ms))
        out["batched_replicas"] = {"replicas": Rr, "days": K, "value": pop.N * Rr * K / (m * 1e-3), "unit": "agent-days/s",
                                   "ms_per_day_all_replicas": m / K,
                                   "what": "%d replicas of the same city (seeds differ) advanced by ONE k_days launch per week; "
                                           "resident, CUDA events" % Rr}
        e.close()
    except engine.DysError as ex:
        out["batched_replicas"] = {"error": str(ex)}
    for key, mode in (("e2e_lockdown", "host"), ("e2e_lockdown_device", "device")):
        try:
            e = engine.Engine(pop, prm, device=local)
            ts = []
            for _ in range(3):
                e.set_state(init["status"], init["t_inf"], init["t_q"], init["t_adm"], init["src"])
                m = EpidemicModel(None, None, None, None, scenario=["GRADUAL", "ALL"], params=prm, population=pop, compact=True,
                                  backend=lambda p_, q_: e, choice_seed=7, lockdown=mode)
                m.run(max_days=W, finalize=False)
                b0 = e.h2d_bytes()
                t = time.perf_counter()
                m.run(max_days=W + K, finalize=False)
                ts.append((time.perf_counter() - t) / max(1, m.day - W))
                h2d = (e.h2d_bytes() - b0) / max(1, m.day - W)
                closed = max(s["Closed buildings"] for s in m.outputs["Results"]["Stats"].values())
            out[key] = {"scenario": "GRADUAL,ALL", "policy": mode, "days": K, "value": pop.N / float(np.median(ts)), "unit": "agent-days/s",
                        "ms_per_step": float(np.median(ts)) * 1e3, "h2d_bytes_per_step": h2d,
                        "max_closed_buildings": int(closed), "window_days": m.window,
                        "api": "EpidemicModel(compact=True, scenario=['GRADUAL','ALL'], lockdown='%s').run()" % mode,
                        "what": ("phase H on the host with the reference's np.random.choice: 4-day windows, the flags travel"
                                 if mode == "host" else
                                 "phase H inside the day kernel (dys_set_lockdown; keyed half closures): week-long windows, no "
                                 "flags travel")}
            e.close()
        except engine.DysError as ex:
            out[key] = {"error": str(ex)}
    try:
        out["network_construction"] = network_extra(local)
    except Exception as ex:                       # noqa: BLE001  (a side measurement must never cost the bench line)
        out["network_construction"] = {"error": repr(ex)}
    return out


def network_extra(local):
    """SURVEY.md 8f-2 on the full shipped city (the committed config-1 inputs: 22 493 agents): the agent-agent edges of
    communities.create_network and the all-pairs shortest paths of contagious3.py:124-130, host arrays in, host arrays out.
    The edge list is checked against the reference's own graph (stored with the fixture) before anything is timed."""
    from dysturbd_b200 import network
    path = os.path.join(ROOT, "tests", "golden", "config1", "config1_inputs.npz")
    if not os.path.exists(path):
        return {"error": "tests/golden/config1/config1_inputs.npz is missing"}
    z = np.load(path)
    n = len(z["agents"])
    cols = network.agent_columns(z["agents"], z["households"], z["build"])
    indptr, col, prob = network.similarity_edges(cols, device=local)
    same = bool(indptr[-1] == len(z["edge_row"]) and np.array_equal(col, z["edge_col"]) and
                np.array_equal(np.repeat(np.arange(n), np.diff(indptr)), z["edge_row"]))
    t = time.perf_counter()
    network.similarity_edges(cols, device=local)
    t_edges = time.perf_counter() - t
    w = network.edge_weights(prob)
    labels = network.components(indptr, col)
    reach = np.bincount(labels)[labels] - 1
    order = np.argsort(labels, kind="stable").astype(np.int32)
    tot = {"sweeps": 0, "prep_s": 0.0, "relax_s": 0.0, "download_s": 0.0}
    pairs = 0
    t = time.perf_counter()
    for k0 in range(0, n, 2048):
        src = order[k0:k0 + 2048]
        _, c, _, info = network.shortest_paths_csr(indptr, col, w, src, int(reach[src].sum()), device=local, return_info=True)
        pairs += len(c)
        for k in tot:
            tot[k] += info[k]
    t_paths = time.perf_counter() - t
    # the rest of the setup: the whole graph G, the routine kernel's neighbourhood lists by bounded searches, the routines
    chain = {}
    try:
        from dysturbd_b200 import routines
        t = time.perf_counter()
        nodes, g_ptr, g_idx, g_w = network.create_graph(z["agents"], z["households"], z["build"], device=local)
        chain["create_graph_s"] = time.perf_counter() - t
        chain["graph_nodes"], chain["graph_edges"] = int(len(nodes)), int(g_ptr[-1])
        t = time.perf_counter()
        nb = routines.neighbourhoods_from_graph(nodes, g_ptr, g_idx, g_w, z["agents"], z["build"], 1, 6, device=local)
        chain["neighbourhood_lists_s"] = time.perf_counter() - t
        chain["list_entries"] = int(len(nb["out"][1]) + len(nb["inn"][1]) + len(nb["near"][1]))
        t = time.perf_counter()
        r = routines.generate(nb, seed=20201019, device=local)
        chain["routines_s"] = time.perf_counter() - t
        chain["mean_visits"] = float((r >= 0).sum() / len(r))
    except Exception as ex:                       # noqa: BLE001
        chain["error"] = repr(ex)
    ref = json.loads(str(z["timings"])) if "timings" in z.files else {}
    return {"workload": "shipped city, %d agents" % n, "edges": int(indptr[-1]), "edges_equal_reference_graph": same,
            "similarity_edges_s": t_edges, "shortest_paths_s": t_paths, "reachable_pairs": int(pairs), **tot,
            "setup_chain": chain, "reference_setup_total_s": ref.get("network+apsp+routines"),
            "reference_network_time_s": ref.get("network_time"), "reference_interaction_time_s": ref.get("interaction_time"),
            "what": "dys_similarity_edges (sizing + fill) and dys_shortest_paths_csr for every agent, wall clock through the Python "
                    "adapter, host arrays in and out; reference figures: contagious3.py's own timers when the fixture was made"}


def eng_bytes(pop):
    return pop.N * (1 + 2 + 16 + 32 + 4 + 2) + pop.nnz // 4 * 16 + pop.N * 8


def cpu_baseline(pop, prm, budget_s, W, K, threads=None, scenario=("noLockdown",)):
    """the C oracle (kind 'port': the literal numpy reference cannot run at this size -- its interaction_prob is a
    dense N x N float64 matrix, 8 TB at 1 M agents) on the same population, all host threads, bounded sample."""
    from oracle import port
    cores = threads or os.cpu_count() or 1
    o = port.OracleSim(pop, prm, threads=cores)      # (explicit: torchrun exports OMP_NUM_THREADS=1)
    t_all, d = 0.0, 0
    while d < W + K and (d < W + 3 or t_all < budget_s):
        t0 = time.perf_counter()
        o.step(d)
        dt = time.perf_counter() - t0
        if d >= W:
            t_all += dt
        d += 1
    days = d - W
    return {"value": pop.N * days / t_all, "unit": "agent-days/s", "cores": cores, "kind": "port", "days": days,
            "sample": "days %d..%d of the same epidemic (%d of %d timed days, flags all open), C oracle, every per-agent loop "
                      "OpenMP-parallel over %d threads" % (W, d - 1, days, K, cores),
            "seconds": t_all}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = args.gpus
    # the same city as the GPU arm: ALL ranks' agents on the host cores (replicas: one city; the replicas are independent)
    if args.replicas > 1:
        pop, n_total = make_population(args, 0, 1)
    else:
        a2 = argparse.Namespace(**vars(args))
        if args.scaling == "weak":
            a2.communities_per_gpu = args.communities_per_gpu * world
        else:
            a2.scaling, a2.communities_per_gpu = "weak", communities(args, 1)
        pop, n_total = make_population(a2, 0, 1)
    prm = make_params(args)
    W, K = args.warmup, args.steps
    cb = cpu_baseline(pop, prm, 60.0, W, K)
    days = cb["days"]
    line = {
        "impl": "reference", "metric": "agent-days/sec", "value": cb["value"], "unit": "agent-days/s", "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": cb["seconds"] * 1e3 / max(1, days), "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "u8/i32 state, f64 probabilities", "data": "synthetic",
        "config": {"workload": workload_string(args, world, pop.N), "agents_total": n_total if args.replicas == 1 else pop.N * args.replicas * world,
                   "norm_factor": args.norm_factor, "scenario": args.scenario,
                   "note": "contagious3.py itself cannot run this size (dense N x N interaction_prob = 8 TB at 1 M agents); this "
                           "is its sparse C restatement (oracle/dys_oracle.c), bit-identical to the literal loop on the golden "
                           "vectors, on all host cores"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "agent-days/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
