import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import srsran_b200 as b
cfg = int(sys.argv[1]); ncb = int(sys.argv[2]) if len(sys.argv) > 2 else 4736
K = 6144
ctx = b.Context(0)
try:
    ctx.set_option("map_cfg", cfg)  # (an engine option of the experiment builds; the release engine has one geometry)
except Exception:
    pass
llr, _ = bench.make_c1(np.random.default_rng(1), ncb, K)
d_llr = ctx.device_alloc(llr.nbytes); d_out = ctx.device_alloc(ncb * K // 8)
ctx.h2d(d_llr, llr)
for _ in range(2):
    ctx.tdec_batch_device(d_llr, d_out, K, ncb, 3 * K + 12, 16, 4)
print("map_ms", ctx.last_map_ms())
