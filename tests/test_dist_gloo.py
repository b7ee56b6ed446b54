"""world_size-2 CPU test (gloo) of the multi-GPU plumbing in bench.py: rank/shard assignment, barrier, max-over-ranks
time and sum-over-ranks work.  The data path itself has no collective (code blocks are independent)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_aggregation(tmp_path):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "dist_worker.py"), str(tmp_path)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    subprocess.run(cmd, check=True, timeout=240, env=env, cwd=ROOT, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    r = [json.load(open(tmp_path / ("rank%d.json" % i))) for i in range(2)]
    assert [x["world"] for x in r] == [2, 2]
    # whole-job throughput = all ranks' units / the slowest rank's time: (1000+2000)*5 bits / 20 ms
    for x in r:
        assert abs(x["ms"] - 20.0) < 1e-9 and abs(x["total"] - 15000) < 1e-9
        assert abs(x["value"] - 15000 / 20e-3 / 1e6) < 1e-12
    assert r[0]["sample"] != r[1]["sample"]  # independent shards
    # pooled cells: the two ranks' job lists are disjoint, cover all 240 jobs, and a cell never changes rank (HARQ affinity)
    j0, j1 = [tuple(x) for x in r[0]["jobs"]], [tuple(x) for x in r[1]["jobs"]]
    assert len(j0) + len(j1) == 240 and not set(j0) & set(j1) and len(set(j0) | set(j1)) == 240
    assert not {c for c, _, _ in j0} & {c for c, _, _ in j1}


def test_reference_arm_other_ranks_exit_quietly():
    """--impl reference under torchrun: only rank 0 runs; the other ranks exit 0 without output"""
    env = dict(os.environ, WORLD_SIZE="2", RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert p.returncode == 0 and p.stdout.strip() == ""
