cd $GRAFT_REPO_ROOT
timeout 100 python tests/gpu_diag.py e2e 2>&1 | grep "fp16 \|FAILED\|rror" 
timeout 100 python tests/variant_time.py 2>&1 | tail -1
B200SEP_FUSED_TAIL=0 timeout 100 python tests/variant_time.py 2>&1 | tail -1
