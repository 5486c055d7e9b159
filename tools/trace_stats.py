import sys, numpy as np
a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 8).astype(np.int64)
n = len(a); t0 = a[:, 0][a[:, 0] > 0].min()
names = ["claim", "k1_start", "tma_landed", "agg_pub", "pfx_pub", "k2_done", "iter_end", "lb_dist"]
sel = slice(n // 4, 3 * n // 4)
def d(i, j): x = (a[sel, i] - a[sel, j]) / 1e3; return f"mean {x.mean():7.2f} p50 {np.median(x):7.2f} p95 {np.percentile(x,95):7.2f} max {x.max():7.2f} us"
print("tiles", n, "span ms", (a[:, 6].max() - t0) / 1e6)
print("claim -> k1_start   ", d(1, 0))
print("k1_start -> landed  ", d(2, 1))
print("landed -> agg_pub   ", d(3, 2))
print("agg_pub -> pfx_pub  ", d(4, 3))
print("pfx_pub -> iter_end ", d(6, 4))
print("iter_end -> k2_done ", d(5, 6))
print("lb_dist: mean", a[sel, 7].mean(), "p95", np.percentile(a[sel, 7], 95), "max", a[sel, 7].max())
# who is the last aggregate each tile waits for: pfx_pub[t] - max(agg_pub[t-dist..t])
lag = []
for t in range(n // 4, n // 4 + 2000):
    dist = int(a[t, 7]); lo = max(0, t - dist)
    lag.append((a[t, 4] - a[lo:t + 1, 3].max()) / 1e3)
print("pfx_pub - latest needed agg: mean %.2f p95 %.2f us" % (np.mean(lag), np.percentile(lag, 95)))
