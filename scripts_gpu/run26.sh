cd $GRAFT_REPO_ROOT
for i in 1 2 3 4 5 6 7 8; do
B200SEP_SYNC_DEBUG=1 B200SEP_LIB=$PWD/audio-classification_b200/variants/lib_KB_TMEM_ONLY_NOPRO.so timeout 100 python tests/variant_time.py 2>&1 | grep -v "^$" | tail -2 | cut -c1-160
done
