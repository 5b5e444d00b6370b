cd $GRAFT_REPO_ROOT
timeout 100 python tests/single_call_latency.py 2>&1 | tail -4
B200SEP_GRAPH=0 timeout 100 python tests/single_call_latency.py 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 100 python tests/split_time.py 2>&1 | tail -1
