cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_aux.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python tests/sweep_sharded.py 600 2>&1 | tail -1 | cut -c1-330
timeout 600 python tests/sweep_sharded.py 600 --two-step 2>&1 | tail -1 | cut -c1-330
