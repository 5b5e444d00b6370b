cd $GRAFT_REPO_ROOT
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
B200SEP_SYNC_DEBUG=1 B200SEP_LIB=$PWD/audio-classification_b200/variants/lib_KB_TMEM_ONLY_NOPRO.so timeout 100 python tests/variant_time.py > gpurun_out/crash_$i.log 2>&1
if grep -q "failed\|rror" gpurun_out/crash_$i.log; then echo "run $i:"; grep -n "b200sep\] launch\|RuntimeError\|rror" gpurun_out/crash_$i.log | head -5; else rm gpurun_out/crash_$i.log; fi
done
echo done
