cd $GRAFT_REPO_ROOT
B200SEP_LIB=$PWD/audio-classification_b200/variants/lib_KA_NO_BOTH.so timeout 100 python tests/timeline.py 10 > gpurun_out/timeline6.log 2>&1
grep -A 30 "CTA 77" gpurun_out/timeline6.log | grep -A3 "K_A" | cut -c1-1400
