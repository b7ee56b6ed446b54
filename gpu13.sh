cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python __graft_entry__.py --smoke 2>&1 | tail -1
( time timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real; tail -c 2500 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | grep real; tail -c 700 gpurun_out/bench_ref.json
