cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python profiles/prof_c1.py 4736 3
python profiles/prof_c1.py 18944 3
python profiles/prof_c2.py 364 3
