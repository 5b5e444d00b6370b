cd $GRAFT_REPO_ROOT
for i in 1 2 3; do timeout 100 python tests/split_time.py 2>&1 | tail -1; done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
