cd $GRAFT_REPO_ROOT
for v in NO_H1_STORE NO_STATS NO_PRO_IO; do
B200SEP_LIB=$PWD/audio-classification_b200/variants/lib_$v.so timeout 200 python tests/variant_time.py 2>&1 | tail -1
done
timeout 200 python tests/variant_time.py 2>&1 | tail -1
