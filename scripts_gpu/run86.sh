cd $GRAFT_REPO_ROOT
timeout 120 python tests/gpu_diag.py e2e 2>&1 | grep "fp16 \|FAILED\|rror\|abort" | head -3
B200SEP_SPLIT=1 timeout 100 python tests/per_block_times.py 2>&1 | tail -27 | head -5
B200SEP_SPLIT=1 B200SEP_LIB=$PWD/audio-classification_b200/variants/lib_NOSHFL.so timeout 100 python tests/per_block_times.py 2>&1 | tail -27 | head -5
for i in 1 2; do
timeout 100 python tests/split_time.py 2>&1 | tail -1
B200SEP_LIB=$PWD/audio-classification_b200/variants/lib_NOSHFL.so timeout 100 python tests/split_time.py 2>&1 | tail -1
done
