cd $GRAFT_REPO_ROOT
timeout 200 python tests/e2e_breakdown.py 2>&1 | tail -5
