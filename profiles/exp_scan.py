"""A/B of the time-parallel latency kernels (map_scan.cuh) against k_map_lat on one subframe: same bytes, same LLR planes, time.
usage: exp_scan.py [K] [ncb] [half_iterations] [amp]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import srsran_b200 as b
K = int(sys.argv[1]) if len(sys.argv) > 1 else 6144
ncb = int(sys.argv[2]) if len(sys.argv) > 2 else 13
nit = int(sys.argv[3]) if len(sys.argv) > 3 else 4
ctx = b.Context(0)
llr, _ = bench.make_c1(np.random.default_rng(1), ncb, K)
if len(sys.argv) > 4:
    llr = np.clip(llr.astype(np.int32) * int(sys.argv[4]) // 100, -32767, 32767).astype(np.int16)
res = {}
for scan in (0, 1, 2):
    ctx.set_option("scan", int(scan > 0))
    ctx.set_option("scan_launch", int(scan == 1))
    ctx.set_option("scan_fused", int(scan == 2))
    out = ctx.tdec_batch(llr, K, nit)
    planes = [np.array(ctx.debug_read_plane(i, pl, K)) for i in range(min(ncb, 4)) for pl in (2, 3)]
    lat = []
    for _ in range(60):
        t0 = time.perf_counter()
        out = ctx.tdec_batch(llr, K, nit)
        lat.append(1e6 * (time.perf_counter() - t0))
    lat = np.sort(lat[10:])
    res[scan] = (out, planes)
    print("scan", scan, "p50_us %.1f" % lat[len(lat) // 2], "map_ms %.4f" % ctx.last_map_ms(), "gpu_ms %.4f" % ctx.last_gpu_ms(), "launches", ctx.last_launches(), "replayed", ctx.last_replayed())
same = all((a == c).all() for a, c in zip(res[0][0], res[1][0])) and all((a == c).all() for a, c in zip(res[0][0], res[2][0]))
same_pl = all((a == c).all() for a, c in zip(res[0][1], res[1][1])) and all((a == c).all() for a, c in zip(res[0][1], res[2][1]))
print("K", K, "ncb", ncb, "nit", nit, "bytes equal", same, "planes equal", same_pl)
if not (same and same_pl):
    for i, (a, c) in enumerate(zip(res[0][1] + res[0][1], res[1][1] + res[2][1])):
        d = np.nonzero(a != c)[0]
        if len(d):
            print(" plane", i, "first diffs at", d[:10], "count", len(d), a[d[:5]], c[d[:5]])
    sys.exit(1)
