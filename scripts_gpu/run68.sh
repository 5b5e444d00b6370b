cd $GRAFT_REPO_ROOT
for v in T_NOENC T_NOTANH T_NOBOTH; do
B200SEP_SPLIT=1 B200SEP_LIB=$PWD/audio-classification_b200/variants/lib_$v.so timeout 100 python tests/variant_time.py 2>&1 | tail -1 | cut -c1-200
done
