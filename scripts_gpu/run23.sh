cd $GRAFT_REPO_ROOT
timeout 100 python tests/gpu_diag.py e2e 2>&1 | grep "fp16 \|FAILED\|rror" 
timeout 100 python tests/variant_time.py 2>&1 | tail -1
B200SEP_LIB=$PWD/audio-classification_b200/variants/lib_KB_TMA_STORE.so timeout 100 python tests/variant_time.py 2>&1 | tail -1
