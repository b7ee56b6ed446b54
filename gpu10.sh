cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python profiles/prof_c1.py 4736 2 > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c1.csv python profiles/prof_c1.py 4736 2 > gpurun_out/ncu0.log 2>&1
cat gpurun_out/prof_plain.log
timeout 900 python bench.py --steps 10 --warmup 3 --ncb 18944 --no-cpu > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo "rc=$?"; tail -c 1500 gpurun_out/bench_c1.json; tail -5 gpurun_out/bench_c1.err
timeout 600 python bench.py --steps 5 --warmup 3 --workload c2 --ntb 1000 --no-cpu > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "rc=$?"; tail -c 1500 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
