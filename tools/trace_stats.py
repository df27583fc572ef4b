#!/usr/bin/env python
"""Developer tool: summarises a DYS_TRACE file (per-block globaltimer stamps of every k_days window).
columns: day block t0..t11; t0 = day start, t1 = after scan push, t2 = agent start, t3 = agent(+push) end, t4 = after the
grid barrier, t11 = service block done (fold + plan)."""
import sys
import numpy as np

rows = np.loadtxt(sys.argv[1], dtype=np.int64)
days = sorted(set(rows[:, 0]))
sel = [int(x) for x in sys.argv[2:]] or days[:: max(1, len(days) // 14)]
print("day | agent phase per block: p10   p50   p90   p99   max | span  | service block | day-to-day")
starts = {}
for d in days:
    r = rows[rows[:, 0] == d]
    G = int(r[:, 1].max()) + 1
    r = r[-G:]                      # last occurrence
    starts[d] = r[1:, 2 + 2].min()
    if d not in sel:
        continue
    w = r[1:]
    t = w[:, 2:].astype(float) / 1e3
    dur = t[:, 3] - t[:, 2]
    span = t[:, 3].max() - t[:, 2].min()
    svc = (r[0, 2 + 11] - r[0, 2 + 4]) / 1e3 if r[0, 2 + 11] > 0 else 0.0
    nxt = starts.get(d + 1)
    print("%4d | %26.1f %5.1f %5.1f %5.1f %5.1f | %5.1f | %8.1f" % ((d,) + tuple(np.percentile(dur, [10, 50, 90, 99, 100])) + (span, svc)))
ds = sorted(starts)
gaps = [(starts[b] - starts[a]) / 1e3 for a, b in zip(ds, ds[1:]) if b == a + 1 and starts[b] > starts[a]]
if gaps:
    print("median day-to-day %.2f us over %d consecutive pairs" % (float(np.median(gaps)), len(gaps)))
