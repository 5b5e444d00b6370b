cd $GRAFT_REPO_ROOT
export B200SEP_LIB=$PWD/audio-classification_b200/variants/lib_KA_PAIR.so
B200SEP_SPLIT=1 timeout 100 python tests/timeline.py 10 > gpurun_out/timeline9.log 2>&1
grep -A 22 "CTA 0" gpurun_out/timeline9.log | grep -A5 "K_A" | cut -c1-1500
