cd $GRAFT_REPO_ROOT
for i in 1 2; do
B200SEP_LIB=$PWD/audio-classification_b200/variants/lib_PREV.so timeout 100 python tests/split_time.py 2>&1 | tail -1
timeout 100 python tests/split_time.py 2>&1 | tail -1
B200SEP_KA_ORDER=0 timeout 100 python tests/split_time.py 2>&1 | tail -1
done
