import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import srsran_b200 as b
K, ncb, nit = 6144, 13, 4
ctx = b.Context(0)
llr, _ = bench.make_c1(np.random.default_rng(1), ncb, K)
for _ in range(3):
    out = ctx.tdec_batch(llr, K, nit)
print("map_ms", ctx.last_map_ms(), "gpu_ms", ctx.last_gpu_ms())
