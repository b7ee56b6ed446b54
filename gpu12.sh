cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python profiles/prof_c2.py 364 2
timeout 600 python bench.py --steps 5 --warmup 3 --workload c2 --ntb 1000 --no-cpu > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "rc=$?"; tail -c 900 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
timeout 600 python bench.py --steps 5 --warmup 3 --workload c4 --ntb 1000 --no-cpu > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "rc=$?"; tail -c 900 gpurun_out/bench_c4.json; tail -5 gpurun_out/bench_c4.err
