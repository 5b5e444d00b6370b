cd $GRAFT_REPO_ROOT
timeout 100 python tests/per_block_times.py 2>&1 | tail -28
