#!/usr/bin/env python
"""Developer tool: summarises a DYS_TRACE file (per-block globaltimer stamps of every k_days window).
columns: day block t0..t7 = day start, after scan push, agent start, agent(+push) end, after barrier,
after sweep, after settle, after local push"""
import sys
import numpy as np

rows = np.loadtxt(sys.argv[1], dtype=np.int64)
sel = [int(x) for x in sys.argv[2:]] or sorted(set(rows[:, 0]))[:: max(1, len(set(rows[:, 0])) // 12)]
seen = set()
print("day   sweep settle  push | block span (first start -> last done) | barrier wait b0 | day-to-day")
starts = {}
for d in sorted(set(rows[:, 0])):
    r = rows[rows[:, 0] == d]
    G = int(r[:, 1].max()) + 1
    r = r[:G]                       # first occurrence (the resident pass)
    starts[d] = r[0, 2]
    if d not in sel:
        continue
    w = r[1:]                       # working blocks (block 0 only folds and plans)
    t = w[:, 2:].astype(float) / 1e3
    ok = t[:, 5] > 0
    sweep = (t[ok, 5] - t[ok, 2]).mean() if ok.any() else 0
    settle = (t[ok, 6] - t[ok, 5]).mean() if ok.any() else 0
    push = (t[ok, 7] - t[ok, 6]).mean() if ok.any() else 0
    span = t[:, 3].max() - t[:, 2].min()
    bw = t[0, 4] - t[0, 3]          # the first working block's wait at the barrier
    nxt = (starts.get(d + 1, 0))
    print("%4d %6.2f %6.2f %6.2f | %6.2f  (agent phase p50 %.2f p99 %.2f max %.2f) | %6.2f" %
          (d, sweep, settle, push, span, np.percentile(t[:, 3] - t[:, 2], 50), np.percentile(t[:, 3] - t[:, 2], 99),
           (t[:, 3] - t[:, 2]).max(), bw))
ds = sorted(starts)
gaps = [(starts[b] - starts[a]) / 1e3 for a, b in zip(ds, ds[1:]) if b == a + 1 and starts[b] > starts[a]]
print("median day-to-day %.2f us over %d consecutive pairs" % (float(np.median(gaps)), len(gaps)))
