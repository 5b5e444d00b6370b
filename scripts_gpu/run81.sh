cd $GRAFT_REPO_ROOT
B200SEP_SPLIT=1 timeout 100 python tests/cta_lifetimes.py 10 2>&1 | tail -5
B200SEP_SPLIT=1 timeout 100 python tests/per_block_times.py 2>&1 | tail -27 | sed -n 4,5p
