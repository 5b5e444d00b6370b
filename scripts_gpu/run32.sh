cd $GRAFT_REPO_ROOT
timeout 100 python tests/gpu_diag.py e2e 2>&1 | grep "fp16 \|FAILED\|rror" 
timeout 100 python tests/per_block_times.py 2>&1 | tail -27 | head -9
timeout 100 python tests/variant_time.py 2>&1 | tail -1
