"""C3-shaped run for profiling: transport blocks over all 188 LTE sizes, device-resident.  usage: prof_c3.py [per_k] [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import srsran_b200 as b  # noqa: E402

per_k = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
specs = bench.c3_specs(per_k)
ctx = b.Context(0)
flat = np.zeros(bench.c3_elems(specs), np.int16)
off = bench.make_c3(np.random.default_rng(1), specs, flat)
ntb = len(specs)
ostr = [(s[0] // 8 + 6 + 15) // 16 * 16 for s in specs]
ooff = np.concatenate([[0], np.cumsum(ostr)[:-1]])
d_llr = ctx.device_alloc(flat.nbytes)
d_out = ctx.device_alloc(int(sum(ostr)))
ctx.h2d(d_llr, flat)
t = b.make_tbs(ntb)
for i in range(ntb):
    t[i].e_bits, t[i].nof_e_bits, t[i].tbs, t[i].Qm, t[i].rv, t[i].softbuffer, t[i].data = d_llr + int(off[i]) * 2, specs[i][2], specs[i][0], specs[i][1], 0, None, d_out + int(ooff[i])
for _ in range(reps):
    ctx.decode_tbs(t, False, 8, flags=b.IN_DEVICE | b.OUT_DEVICE)
print("ntb", ntb, "gpu_ms", ctx.last_gpu_ms(), "map_ms", ctx.last_map_ms(), "launches", ctx.last_launches(), "avg_it", np.mean([t[i].avg_iterations for i in range(ntb)]),
      "ok", sum(1 for i in range(ntb) if t[i].ret == 0))
