cd $GRAFT_REPO_ROOT
for i in 1 2 3; do timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2; done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for i in 1 2 3 4 5 6; do timeout 100 python tests/split_time.py 2>&1 | tail -1; done
for i in 1 2 3; do timeout 100 python tests/e2e_profile.py 2>&1 | tail -2; done
