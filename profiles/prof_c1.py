"""Small C1-shaped run for profiling: NCB code blocks of K=6144, 4 half-iterations, device-resident input."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import srsran_b200 as b  # noqa: E402

ncb = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
K = 6144
ctx = b.Context(0)
llr, _ = bench.make_c1(np.random.default_rng(1), ncb, K)
d_llr = ctx.device_alloc(llr.nbytes)
d_out = ctx.device_alloc(ncb * K // 8)
ctx.h2d(d_llr, llr)
for _ in range(reps):
    ctx.tdec_batch_device(d_llr, d_out, K, ncb, 3 * K + 12, 16, 4)
print("gpu_ms", ctx.last_gpu_ms(), "map_ms", ctx.last_map_ms(), "replayed", ctx.last_replayed())
