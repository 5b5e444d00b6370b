cd $GRAFT_REPO_ROOT
timeout 100 python tests/e2e_profile.py 2>&1 | tail -6
