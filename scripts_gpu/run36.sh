cd $GRAFT_REPO_ROOT
timeout 100 python tests/gpu_diag.py e2e 2>&1 | grep "fp16 \|FAILED\|rror" | head
for i in 1 2 3; do timeout 100 python tests/variant_time.py 2>&1 | tail -1 | cut -c1-200; done
timeout 100 python tests/per_block_times.py 2>&1 | tail -27 | head -9
