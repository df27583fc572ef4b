import sys, numpy as np
rows = np.loadtxt(sys.argv[1], dtype=np.uint64)
for d in (40, 64, 80):
    r = rows[rows[:, 0] == d]
    G = int(r[:, 1].max()) + 1
    svc = r[:G][0]
    r = r[:G][1:]
    t = r[:, 2:].astype(np.float64)
    dur = (t[:, 3] - t[:, 2]) / 1e3
    nl = (r[:, 3] >> np.uint64(32)).astype(float); nf = (r[:, 3] & np.uint64(0xFFFFFFFF)).astype(float)
    start = (t[:, 2] - t[:, 2].min()) / 1e3
    print('day', d, 'dur p50 %.1f max %.1f | list mean %.0f sd %.0f max %.0f | front mean %.0f sd %.0f max %.0f' % (np.median(dur), dur.max(), nl.mean(), nl.std(), nl.max(), nf.mean(), nf.std(), nf.max()))
    print('   corr(dur,list) %.2f corr(dur,front) %.2f corr(dur, blockIdx) %.2f corr(dur, blk%%148) %.2f  start spread p50 %.1f max %.1f' % (
        np.corrcoef(dur, nl)[0, 1], np.corrcoef(dur, nf)[0, 1], np.corrcoef(dur, np.arange(len(dur)))[0, 1], np.corrcoef(dur, np.arange(len(dur)) % 148)[0, 1], np.median(start), start.max()))
    A = np.stack([np.ones_like(nl), nl, nf], 1)
    coef, *_ = np.linalg.lstsq(A, dur, rcond=None)
    res = dur - A @ coef
    print('   fit dur = %.2f + %.4f*list + %.4f*front ; residual sd %.2f' % (coef[0], coef[1], coef[2], res.std()))
    sw = (t[:, 5] - t[:, 2]) / 1e3; se = (t[:, 6] - t[:, 5]) / 1e3; pu = (t[:, 7] - t[:, 6]) / 1e3
    print('   phases p50/max: sweep %.1f/%.1f settle %.1f/%.1f push %.1f/%.1f' % (np.median(sw), sw.max(), np.median(se), se.max(), np.median(pu), pu.max()))
    if t.shape[1] > 10:
        ld = (t[:, 8] - t[:, 5]) / 1e3; st_ = (t[:, 9] - t[:, 8]) / 1e3; ag = (t[:, 10] - t[:, 9]) / 1e3; rest = (t[:, 6] - t[:, 10]) / 1e3
        print('   settle detail (warp 0, first round) p50/max: loads %.1f/%.1f rules+raise %.1f/%.1f counters+events %.1f/%.1f rest-of-phase %.1f/%.1f' % (
            np.median(ld), ld.max(), np.median(st_), st_.max(), np.median(ag), ag.max(), np.median(rest), rest.max()))
    if len(svc) > 13:
        f = lambda i: float(svc[2 + i]) / 1e3
        print('   service block after the barrier: fold %.1f, stage loads %.1f, sums+scan %.1f, walk %.1f us' % (f(8) - f(4), f(9) - f(8), f(10) - f(9), f(11) - f(10)))
    idx = np.argsort(-dur)[:8]
    print('   slowest blocks', [(int(i), round(float(dur[i]), 1), int(nl[i]), int(nf[i])) for i in idx])
