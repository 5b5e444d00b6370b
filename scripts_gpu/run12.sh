cd $GRAFT_REPO_ROOT
B200SEP_LIB=$PWD/audio-classification_b200/variants/lib_NO_COEF_CVT.so timeout 100 python tests/variant_time.py 2>&1 | tail -1
timeout 100 python tests/variant_time.py 2>&1 | tail -1
timeout 100 python tests/gpu_diag.py e2e 2>&1 | grep "fp16 " 
