cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 20 --no-cpu > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo "rc=$?"; tail -c 1800 gpurun_out/bench_c1.json; tail -5 gpurun_out/bench_c1.err
timeout 600 python bench.py --steps 10 --warmup 3 --workload c2 --ntb 1000 --no-cpu > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "rc=$?"; tail -c 1800 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
