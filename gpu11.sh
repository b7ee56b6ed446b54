cd $GRAFT_REPO_ROOT
python profiles/prof_c2.py 364 2 > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c2.csv python profiles/prof_c2.py 364 2 > gpurun_out/ncu0.log 2>&1
cat gpurun_out/prof_plain.log
