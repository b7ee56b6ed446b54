"""Experiment: time the MAP kernel variants (engine option map_cfg) on a C1-shaped batch, check outputs agree."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import srsran_b200 as b

ncb = int(sys.argv[1]) if len(sys.argv) > 1 else 18944
cfgs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2, 3, 4]
K = 6144
ctx = b.Context(0)
llr, _ = bench.make_c1(np.random.default_rng(1), ncb, K)
d_llr = ctx.device_alloc(llr.nbytes)
d_out = ctx.device_alloc(ncb * K // 8)
ctx.h2d(d_llr, llr)
ref = None
for cfg in cfgs:
    try:
        ctx.set_option("map_cfg", cfg)  # (an engine option of the experiment builds; the release engine has one geometry)
    except Exception:
        pass
    best = 1e9
    for _ in range(4):
        ctx.tdec_batch_device(d_llr, d_out, K, ncb, 3 * K + 12, 16, 4)
        best = min(best, ctx.last_map_ms())
    out = np.zeros(ncb * K // 8, np.uint8)
    ctx.d2h(out, d_out)
    if ref is None:
        ref = out
    same = bool((out == ref).all())
    print("cfg %d: map_ms %.3f gpu_ms %.3f  -> %.1f Gbit/s (MAP only)  replayed %d same=%s" % (
        cfg, best, ctx.last_gpu_ms(), ncb * K / best / 1e6, ctx.last_replayed(), same), flush=True)
