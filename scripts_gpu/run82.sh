cd $GRAFT_REPO_ROOT
timeout 120 python tests/gpu_diag.py e2e 2>&1 | grep "n_src=3 fp16 \|FAILED\|rror\|abort" | head -3
B200SEP_SPLIT=1 timeout 100 python tests/cta_lifetimes.py 10 2>&1 | grep "K_A"
B200SEP_SPLIT=1 timeout 100 python tests/variant_time.py 2>&1 | tail -1 | cut -c1-200
B200SEP_SPLIT=1 B200SEP_KA_ORDER=0 timeout 100 python tests/variant_time.py 2>&1 | tail -1 | cut -c1-200
for i in 1 2; do timeout 100 python tests/split_time.py 2>&1 | tail -1; done
B200SEP_KA_ORDER=0 timeout 100 python tests/split_time.py 2>&1 | tail -1
