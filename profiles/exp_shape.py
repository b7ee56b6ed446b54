"""code-block batch of a chosen shape through the fused kernel: exp_shape.py K ncb half_iterations [reps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import srsran_b200 as b
K = int(sys.argv[1]); ncb = int(sys.argv[2]); nit = int(sys.argv[3]); reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
ctx = b.Context(0)
llr, _ = bench.make_c1(np.random.default_rng(1), ncb, K)
d_llr = ctx.device_alloc(llr.nbytes); d_out = ctx.device_alloc(ncb * K // 8)
ctx.h2d(d_llr, llr)
for _ in range(reps):
    ctx.tdec_batch_device(d_llr, d_out, K, ncb, 3 * K + 12, 16, nit)
ms = ctx.last_map_ms()
print("K", K, "ncb", ncb, "nit", nit, "map_ms", ms, "ns per block-half-iteration", 1e6 * ms / (ncb * nit), "gpu_ms", ctx.last_gpu_ms())
