cd $GRAFT_REPO_ROOT
for i in 1 2 3 4; do
B200SEP_LIB=$PWD/audio-classification_b200/variants/lib_KB_TMEM_ONLY_NOPRO.so timeout 100 python tests/variant_time.py 2>&1 | tail -1 | cut -c1-100
done
echo "--- PDL on, default lib"
for i in 1 2 3 4 5 6 7 8; do
B200SEP_PDL=1 timeout 100 python tests/variant_time.py 2>&1 | tail -1 | cut -c1-40
done
