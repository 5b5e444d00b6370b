cd $GRAFT_REPO_ROOT
for v in KA_NO_COMPUTE KA_NO_EPI_STORE KA_NO_BOTH; do
B200SEP_LIB=$PWD/audio-classification_b200/variants/lib_$v.so timeout 100 python tests/variant_time.py 2>&1 | tail -1 | cut -c1-130
done
