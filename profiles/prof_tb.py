"""C2 / C4-shaped run for profiling: NTB transport blocks through the batched decode_tb entry, device-resident.
usage: prof_tb.py [ntb] [reps] [c2|c4]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import srsran_b200 as b  # noqa: E402

ntb = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
wl = sys.argv[3] if len(sys.argv) > 3 else "c2"
cfg = dict(bench.TB_CFG[wl])
if os.environ.get("PROF_SIGMA"):
    cfg["sigma"] = float(os.environ["PROF_SIGMA"])
if os.environ.get("PROF_MAX_ITER"):
    cfg["max_iter"] = int(os.environ["PROF_MAX_ITER"])
tbs, Qm, G, dt = cfg["tbs"], cfg["Qm"], cfg["G"], cfg["dtype"]
ctx = b.Context(0)
llr, _ = bench.make_tb(np.random.default_rng(1), ntb, tbs, Qm, G, dt, cfg["amp"], cfg["sigma"])
esz = np.dtype(dt).itemsize
ostride = (tbs // 8 + 6 + 15) // 16 * 16
d_llr = ctx.device_alloc(llr.nbytes)
d_out = ctx.device_alloc(ntb * ostride)
ctx.h2d(d_llr, llr)
t = b.make_tbs(ntb)
for i in range(ntb):
    t[i].e_bits, t[i].nof_e_bits, t[i].tbs, t[i].Qm, t[i].rv, t[i].softbuffer, t[i].data = d_llr + i * G * esz, G, tbs, Qm, 0, None, d_out + i * ostride
for _ in range(reps):
    ctx.decode_tbs(t, dt == np.int8, cfg["max_iter"], flags=b.IN_DEVICE | b.OUT_DEVICE)
print("gpu_ms", ctx.last_gpu_ms(), "map_ms", ctx.last_map_ms(), "replayed", ctx.last_replayed(), "avg_it", np.mean([t[i].avg_iterations for i in range(ntb)]),
      "ok", sum(1 for i in range(ntb) if t[i].ret == 0), "half_iter", ctx.last_half_iterations() if hasattr(ctx, "last_half_iterations") else None)
