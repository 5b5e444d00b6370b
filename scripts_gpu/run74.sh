cd $GRAFT_REPO_ROOT
for v in KB_D8 KB_D10; do
export B200SEP_LIB=$PWD/audio-classification_b200/variants/lib_$v.so
timeout 120 python tests/gpu_diag.py e2e 2>&1 | grep "n_src=3 fp16 \|FAILED\|rror\|abort" | head -3
B200SEP_SPLIT=1 timeout 100 python tests/variant_time.py 2>&1 | tail -1 | cut -c1-200
timeout 100 python tests/split_time.py 2>&1 | tail -1
done
