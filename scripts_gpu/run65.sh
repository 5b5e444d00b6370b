cd $GRAFT_REPO_ROOT
B200SEP_PDL=1 timeout 100 python tests/single_call_latency.py 2>&1 | tail -4
B200SEP_PDL=1 B200SEP_GRAPH=0 timeout 100 python tests/single_call_latency.py 2>&1 | tail -2
B200SEP_PDL=1 timeout 100 python tests/gpu_diag.py e2e 2>&1 | grep "n_src=3 fp16 \|FAILED\|rror\|abort" | head -3
B200SEP_PDL=1 timeout 100 python tests/split_time.py 2>&1 | tail -1
