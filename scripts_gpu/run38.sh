cd $GRAFT_REPO_ROOT
timeout 200 python tests/wave_sweep.py 2>&1 | tail -10
