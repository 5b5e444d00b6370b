cd $GRAFT_REPO_ROOT
fails=0
for i in 1 2 3 4 5 6 7 8 9 10 11 12 13 14 15 16; do
B200SEP_LIB=$PWD/audio-classification_b200/variants/lib_KB_TMEM_ONLY_NOPRO.so timeout 100 python tests/variant_time.py > gpurun_out/c.log 2>&1
if grep -q "failed\|rror" gpurun_out/c.log; then fails=$((fails+1)); fi
done
echo "fast variant: $fails / 16 failed"
fails=0
for i in 1 2 3 4 5 6 7 8 9 10; do
B200SEP_PDL=1 timeout 100 python tests/variant_time.py > gpurun_out/c.log 2>&1
if grep -q "failed\|rror" gpurun_out/c.log; then fails=$((fails+1)); fi
done
echo "default lib, PDL on: $fails / 10 failed"; tail -1 gpurun_out/c.log | cut -c1-150
timeout 100 python tests/variant_time.py 2>&1 | tail -1
