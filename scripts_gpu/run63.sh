cd $GRAFT_REPO_ROOT
timeout 100 python tests/single_call_latency.py 2>&1 | tail -4
