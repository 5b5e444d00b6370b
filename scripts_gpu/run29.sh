cd $GRAFT_REPO_ROOT
for v in NO_PRO_IO KB_NO_L2PF; do
B200SEP_LIB=$PWD/audio-classification_b200/variants/lib_$v.so timeout 100 python tests/variant_time.py 2>&1 | tail -1 | cut -c1-130
done
