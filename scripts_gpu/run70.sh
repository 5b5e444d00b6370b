cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_aux.py -m gpu -x -q -s -k "separate_and_score" 2>&1 | tail -6
