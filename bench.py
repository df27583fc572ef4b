#!/usr/bin/env python
"""bench.py -- agent-days/s of the per-day epidemic step (BASELINE.json metric) on N B200s of one node.

  python bench.py --gpus 1 --steps K --warmup W            (N=1: the synthetic ~1 M-agent city, BASELINE.json configs[1])
  torchrun ... bench.py --gpus N --steps K --warmup W      (N>1: ONE coupled city of N x ~1 M agents, sharded by
                                                            community, per-day NCCL exchange; weak scaling)
  python bench.py --impl reference ...                     (the CPU path on the host cores, same config and metric)

A step is one simulated day (one pass of contagious3.py:176-366) over the whole population.  Prints ONE JSON line.
`value`  : K days back to back with the population resident in HBM, timed with CUDA events on the library's stream.
`e2e`    : the same K days through the C ABI the way the reference-facing host loop uses it: every day the open
           flags come from pinned host memory and the day's outputs (stats, per-area histograms, per-building counts,
           infection events) are read back to host memory.
`roofline`: k_days (the fused day kernel) -- algorithmic bytes per simulated day from the day's own work counters
           (formula in DESIGN.md) over its mean duration (CUDA events around the resident pass), against
           MEASURED_PEAKS.json; `survey_8d` evaluates SURVEY.md 8(d)'s formula with the same counters.
`cpu_baseline`: the C oracle (OpenMP) on the same population, a bounded sample of days, rank 0 / N=1 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=365)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--communities-per-gpu", type=int, default=89, help="89 x 11 264 = 1 002 496 agents per GPU")
    ap.add_argument("--norm-factor", type=float, default=6.0,
                    help="with the degree-16.7 synthetic network 6.0 gives a reference-like epidemic: peak near day 70, "
                         "38 %% of the population infected, extinct near day 280")
    ap.add_argument("--seeds-per-22493", type=int, default=20)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cpu-baseline-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N>1: fused P2P stores of the snapshot bytes from k_agent (peer) or a per-day NCCL all-gather")
    return ap.parse_args()


def make_population(args, rank, world):
    from dysturbd_b200.synth import COMMUNITY, synth_population
    cpg = args.communities_per_gpu
    n_total = COMMUNITY * cpg * world
    comm = None if world == 1 else (rank * cpg, (rank + 1) * cpg)
    pop = synth_population(n_total, seed=args.seed, communities=comm, seeds_per_22493=args.seeds_per_22493)
    return pop, n_total


def make_params(args):
    from dysturbd_b200.population import EpiParams
    return EpiParams(norm_factor=args.norm_factor, seed=20201019)


class ClockSampler:
    """SM clocks and throttle reasons DURING the timed regions.  The timed regions are tens of milliseconds long, so
    the nvidia-smi loop of B200_PROFILING.md (200 ms period) would miss them: the same NVML counters are polled from a
    thread every millisecond instead."""

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
            time.sleep(0.001)

    def sample_now(self):
        """one synchronous sample (a timed region can be shorter than the polling period)"""
        nv = self.nv
        if nv is None:
            return
        try:
            sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
            try:
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            self.rows.append((sm, rs))
        except Exception:
            pass

    def stop(self):
        self._stop = True

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        nv = self.nv
        names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4,
                 "hw_power_brake": 0x80}
        seen = set()
        for _, rs in self.rows:
            for n, bit in names.items():
                if rs & bit:
                    seen.add(n)
        return {"sm_mhz": float(np.median([r[0] for r in self.rows])), "sm_max_mhz": float(self.max_sm), "reasons": sorted(seen),
                "samples": len(self.rows)}


def kernel_bytes(N, st, prev, window=7):
    """Algorithmic bytes one simulated day must move in this design (DESIGN.md 'bytes per day'), from the day's own
    counters: each needed datum once, at its stored width.  `st` = today's stats block, `prev` = yesterday's (its
    end-of-day status counts are today's day-start counts)."""
    from dysturbd_b200.engine import STAT
    g = lambda k: int(st[STAT[k]])
    S0, I0, QH0, QI0, D0, H0, R0, X0 = (int(x) for x in prev[20:28])
    rows = g("inf0")
    # push: the row extent of every infectious agent, its kept network entries (16-byte records), a mailbox + attention
    # byte per success; the snapshot bytes are only scanned on the first day of a window
    push = rows * 16 + g("edges_scanned") * 16 + g("hits") * (4 + 4 + 1) + N // window
    # agent: status + attention byte of everybody; the 16-byte record half of every agent that is not idle; state writes
    listed = I0 + QI0 + D0 + H0 + X0 + QH0 + g("hits") + g("daily_quar")
    agent = N * (1 + 1) + listed * 16
    agent += g("changed") * (1 + 2) + g("n_events") * (8 + 4) + g("daily_quar") * 2
    agent += int(st[32:53].sum()) * 4 + (I0 + QI0 + D0) * 4      # per-area histogram bin, per-building count
    agent += 2 * N // window                                    # tomorrow's snapshot bytes: last day of a window only
    return {"k_push": push, "k_agent": agent}


def survey_8d_bytes(N, st, prev, mean_degree):
    """SURVEY.md section 8(d)'s per-day byte formula (a straightforward pull over compact SoA columns), evaluated with
    the same counters -- what the PROBLEM would cost an implementation that re-reads routines and gates rows every day."""
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

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus must equal WORLD_SIZE")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pop, n_total = make_population(args, rank, world)
    prm = make_params(args)
    W, K = args.warmup, args.steps
    all_out = engine.WANT_SA_HIST | engine.WANT_BUILDING_COUNTS | engine.WANT_EVENTS
    from dysturbd_b200.sharded import ShardedEngine

    def make_engine(flags):
        if world == 1:
            return engine.Engine(pop, prm, device=local, flags=flags)
        return ShardedEngine(pop, prm, device=local, mode=args.exchange, flags=flags)

    def raw(e):                      # the ctypes Engine underneath
        return e if world == 1 else e.sim

    t0 = time.perf_counter()
    eng = make_engine(all_out)
    eng_mode = getattr(eng, "mode", "single")
    raw(eng).sync()
    load_s = time.perf_counter() - t0
    stream = torch.cuda.ExternalStream(raw(eng).cuda_stream(), device=local)
    init = pop.copy_state()

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

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()

    # ---- pass 1: resident throughput (`value`) ----
    def run_resident(first, last):
        # weeks of days per call: dys_step_many runs each window as ONE cooperative launch (k_days)
        d = first
        while d < last:
            n = min(7, last - d)
            eng.step_many(d, n)
            d += n

    with torch.cuda.stream(stream):
        run_resident(0, W)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        clocks.active = True
        ev0.record(stream)
        run_resident(W, W + K)
        clocks.sample_now()          # queued, not finished: the GPU is inside the timed region here
        raw(eng).sync()              # the last days' output blocks are still leaving on the library's copy stream
        ev1.record(stream)
        barrier()
        clocks.active = False
        ms_value = max_over_ranks(ev0.elapsed_time(ev1))
    st_last = raw(eng).stats(W + K - 1)

    # ---- pass 2: end to end through the C ABI with host buffers (`e2e`) ----
    # The host loop of dysturbd_b200/host.py (EpidemicModel.run): days are queued in windows with two windows in flight
    # (week-long for a scenario that closes nothing, as here; (diagnosis + 1) // 2 = 4 days under a lockdown scenario,
    # whose decision lags `diagnosis` days, contagious3.py:306-309).  Every window the open flags of its days are offered
    # from pinned host memory (dys_step_many compares them with the vector in force and only transfers a changed one) and
    # every day's whole output block (stats, per-area histograms, per-building counts, infection events) is read on the
    # host from the pinned ring the library copied it into.
    WIN = 7
    eng.set_state(init["status"], init["t_inf"], init["t_q"], init["t_adm"], init["src"])
    barrier()                        # (peer mode: nobody steps before every rank has reset its flags)
    open_pinned = torch.ones((WIN, pop.B), dtype=torch.uint8).pin_memory()
    h2d = d2h = 0
    stats_days, pending = [], []
    checksum = 0
    side = torch.cuda.Stream(device=local)
    with torch.cuda.stream(stream):
        def consume_day(d, count):
            nonlocal d2h, checksum
            o = eng.outputs(d)
            st = o["stats"].copy()
            # the block is in host memory now (one D2H per day, queued by the library); touch each part
            checksum += int(o["sa_hist"][-1, 0]) + int(o["building_counts"][-1]) + o["n_events"] + (int(o["events"][-1, 0]) if len(o["events"]) else 0)
            if count:
                d2h += st.nbytes + o["sa_hist"].nbytes + o["building_counts"].nbytes + o["events"].nbytes
                stats_days.append(st)
                pending.append(st)
            if world > 1 and pending and (len(pending) == 7 or d == W + K - 1):
                # global stats (R, the lockdown decision) are only needed `diagnosis` = 7 days later
                # (contagious3.py:306-309): one all-reduce per week of stacked 64-integer blocks
                with torch.cuda.stream(side):          # not on the library's stream: the days keep flowing
                    t = torch.from_numpy(np.stack(pending)).cuda()
                    dist.all_reduce(t)
                    g = t.cpu().numpy()
                for k in range(len(pending)):
                    stats_days[len(stats_days) - len(pending) + k] = g[k]
                pending.clear()

        def run_days(first, last, count):
            queued = consumed = first
            while consumed < last:
                while queued < min(last, consumed + 2 * WIN):
                    n = min(WIN, last - queued)
                    eng.step_many(queued, n, open_pinned[:n])
                    queued += n
                for _ in range(min(WIN, queued - consumed)):
                    consume_day(consumed, count)
                    consumed += 1

        run_days(0, W, False)
        barrier()
        clocks.active = True
        t0 = time.perf_counter()
        run_days(W, W + K, True)
        clocks.sample_now()
        barrier()
        s_e2e = max_over_ranks(time.perf_counter() - t0)
        clocks.active = False
    assert np.array_equal(stats_days[-1][:12], st_last[:12]) or world > 1, "e2e replay diverged from the resident pass"

    # ---- pass 3: per-kernel CUDA-event durations for the roofline (same days, same state) ----
    eng.close()
    eng_t = make_engine(all_out | engine.TIME_KERNELS)
    stream_t = torch.cuda.ExternalStream(raw(eng_t).cuda_stream(), device=local)
    with torch.cuda.stream(stream_t):
        for d in range(W):
            eng_t.step(d)
        raw(eng_t).kernel_times()
        prev_w = raw(eng_t).stats(W - 1) if W > 0 else None      # day-start compartment counts of the first timed day
        barrier()
        for d in range(W, W + K):
            eng_t.step(d)
        ms_k, n_k = raw(eng_t).kernel_times()
    prev = prev_w if prev_w is not None and world == 1 else stats_days[0]
    kb = {"k_push": 0, "k_agent": 0}
    survey_bytes = 0
    for st in stats_days:                   # the e2e pass ran the same days from the same state: same counters
        for k, v in kernel_bytes(n_total, st, prev).items():
            kb[k] += v // world             # counters are global after the all-reduce; a rank moves its share
        survey_bytes += survey_8d_bytes(n_total, st, prev, pop.nnz / pop.N) // world
        prev = st
    eng_t.close()
    clocks.stop()

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    per_kernel = {}
    for i, name in enumerate(engine.KERNEL_NAMES):
        if name in kb and n_k[i]:
            ms = ms_k[i] / n_k[i]
            per_kernel[name] = {"ms_per_launch": ms, "algorithmic_bytes_per_launch": kb[name] / K,
                                "achieved_gbs": kb[name] / K / (ms * 1e-3) / 1e9}
    # the product path runs a window of days as ONE cooperative launch (k_days = push_body + agent_body per day): its
    # per-day duration is the resident pass above (CUDA events on the launching stream), its bytes the two bodies' sum
    dom = "k_days (per simulated day: agent pass + push of the block's own frontier, one grid barrier; one cooperative launch per 7 days)"
    k_ms = ms_value / K
    kbytes = sum(kb.values())
    achieved = kbytes / K / (k_ms * 1e-3) / 1e9

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = n_total * K / (ms_value * 1e-3)
    line = {
        "metric": "agent-days/sec", "value": value, "unit": "agent-days/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_value / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/i16 state, f64 probabilities", "data": "synthetic",
        "config": {
            "workload": "synthetic city, %d agents per GPU (%d communities of 11264; ~300k households, ~50k buildings per "
                        "million), communities social network (degree ~16.7), days %d..%d of one epidemic" %
                        (pop.N, args.communities_per_gpu, W, W + K - 1),
            "agents_total": n_total, "edges_per_gpu": pop.nnz, "norm_factor": args.norm_factor,
            "seeds_per_22493": args.seeds_per_22493, "scenario": "noLockdown",
            "parallelism": "1 GPU" if world == 1 else "sharded by community x%d, exchange=%s (peer: snapshot bytes stored into every peer GPU by k_agent over NVLink; nccl: per-day all-gather)" % (world, eng_mode),
            "l2": "no flush: the inputs of a day (%.0f MB of state + network per GPU) are larger than the 126 MB L2 and "
                  "every day rewrites the state" % (eng_bytes(pop) / 1e6),
            "ever_infected_at_end": int(st_last[STAT["ever_infected"]]) if world == 1 else None,
            "load_seconds": load_s,
        },
        "e2e": {"value": n_total * K / s_e2e, "unit": "agent-days/s", "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K,
                "ms_per_step": s_e2e * 1e3 / K,
                "h2d_note": "the only per-day input of this path is the building open-flag vector (%d bytes), offered every "
                            "day from pinned host memory and transferred when it differs from the one in force -- never "
                            "under noLockdown; the population is resident by construction (it is the simulation state)" % pop.B},
        "gpu_launches": -(-K // 7),  # one cooperative k_days launch per 7-day window (k_snapshot only after a load)
        "kernels_ms_per_step": {n: ms_k[i] / max(1, n_k[i]) for i, n in enumerate(engine.KERNEL_NAMES)},
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if peak else None,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one k_days launch / its 7 simulated days, from the
                     # committed ncu --set full capture (profiles/r01k_summary.md: days 66-72, the peak week, whose
                     # algorithmic bytes are ~8 MB per day); not measured live
                     "traffic": (49410560 + 2254080) / 7 if world == 1 and args.communities_per_gpu == 89 else None,
                     "traffic_note": "per simulated day, peak week of the epidemic (ncu capture of 2026-10-20, profiles/r01k_summary.md)",
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 (B200_PROFILING.md)",
                     "algorithmic_bytes_per_launch": kbytes / max(1, K), "ms_per_launch": k_ms,
                     "note": "algorithmic bytes = what THIS design must move per day (bench.py kernel_bytes, DESIGN.md); a "
                             "1 M-agent day is a chain of dependent memory round trips and one grid barrier, not a bandwidth "
                             "problem -- see profiles/ for the per-phase trace",
                     "survey_8d": {"bytes_per_day": survey_bytes / max(1, K),
                                   "achieved": survey_bytes / max(1, K) / (k_ms * 1e-3) / 1e9,
                                   "frac": survey_bytes / max(1, K) / (k_ms * 1e-3) / 1e9 / peak if peak else None,
                                   "what": "SURVEY.md 8(d)'s per-day formula (daily routine rows, gated pull over every "
                                           "interaction_prob non-zero) with the same counters: the bytes the problem costs a "
                                           "straightforward SoA pull, which this design avoids moving"},
                     "per_kernel_standalone": per_kernel},
        "clocks": clocks.summary(),
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(pop, prm, args.cpu_baseline_seconds, W, K)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def eng_bytes(pop):
    return pop.N * (1 + 6 + 4 + 12 + 24 + 1) + pop.nnz * 12 + pop.routine.size * 4


def cpu_baseline(pop, prm, budget_s, W, K, threads=None):
    """the C oracle (kind 'port': the literal numpy reference cannot run at this size -- its interaction_prob is a
    dense N x N float64 matrix, 8 TB at 1 M agents) on the same population, all host threads, bounded sample."""
    from oracle import port
    o = port.OracleSim(pop, prm, threads=threads)
    cores = threads or port.max_threads()
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
            "sample": "days %d..%d of the same epidemic (%d of %d timed days), C oracle with OpenMP" % (W, d - 1, days, K),
            "seconds": t_all}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pop, n_total = make_population(args, 0, 1)
    prm = make_params(args)
    W, K = args.warmup, args.steps
    cb = cpu_baseline(pop, prm, 120.0, W, K)
    days = cb["days"]
    line = {
        "impl": "reference", "metric": "agent-days/sec", "value": cb["value"], "unit": "agent-days/s", "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": cb["seconds"] * 1e3 / max(1, days), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/i32 state, f64 probabilities", "data": "synthetic",
        "config": {"workload": "synthetic city, %d agents (%d communities of 11264), CPU path on the host cores" %
                               (pop.N, args.communities_per_gpu), "norm_factor": args.norm_factor,
                   "note": "contagious3.py itself cannot run this size (dense N x N interaction_prob = 8 TB); this is its "
                           "sparse C restatement, bit-identical to the literal loop on the golden vectors"},
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
