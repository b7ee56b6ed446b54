"""worker of tests/test_dist_gloo.py: exercises bench.py's multi-rank plumbing on the CPU (gloo, world_size 2)"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

rank, world, local, barrier, vmax, vsum = bench.dist_setup(2, backend="gloo")
barrier()
# every rank owns an independent shard (different seed -> different code blocks), no data-path collective
rng = np.random.default_rng(bench.shard_seed(rank))
sample = rng.integers(0, 1 << 30, 4).tolist()
units = 1000 * (rank + 1)  # pretend rank r decoded 1000*(r+1) bits per step
ms = 10.0 * (rank + 1)     # and needed 10*(r+1) ms
value, ms_max, total = bench.aggregate(units, 5, ms, vmax, vsum)
barrier()
# pooled cells (BASELINE config 5): every rank takes the jobs of the cells it owns -- 20 cells x 6 TTIs x (dl, ul)
from srsran_b200.pool import Job, partition  # noqa: E402
jobs = [Job(c, t, k, t % 8, 0, True, 1000, 2, None) for t in range(6) for c in range(20) for k in ("dl", "ul")]
mine = [(j.cell, j.tti, j.kind) for j in partition(jobs, rank, world)]
out = {"rank": rank, "world": world, "value": value, "ms": ms_max, "total": total, "sample": sample, "jobs": mine}
with open(os.path.join(sys.argv[1], "rank%d.json" % rank), "w") as f:
    json.dump(out, f)
import torch.distributed as dist  # noqa: E402
dist.destroy_process_group()
