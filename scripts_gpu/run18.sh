cd $GRAFT_REPO_ROOT
for v in KA_NO_B_TMA KA_HALF_N KA_BOTH; do
B200SEP_LIB=$PWD/audio-classification_b200/variants/lib_$v.so timeout 100 python tests/variant_time.py 2>&1 | tail -1
done
timeout 100 python tests/variant_time.py 2>&1 | tail -1
