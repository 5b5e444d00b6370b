cd $GRAFT_REPO_ROOT
timeout 120 python tests/gpu_diag.py e2e 2>&1 | grep "fp16 \|fp32 \|FAILED\|rror\|abort" | head -6
B200SEP_SPLIT=1 timeout 100 python tests/variant_time.py 2>&1 | tail -1 | cut -c1-200
for i in 1 2; do timeout 100 python tests/split_time.py 2>&1 | tail -1; done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
