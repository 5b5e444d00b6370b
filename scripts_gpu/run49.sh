cd $GRAFT_REPO_ROOT
B200SEP_SPLIT=1 timeout 100 python tests/timeline.py 10 > gpurun_out/timeline8.log 2>&1
grep -A 12 "CTA 77" gpurun_out/timeline8.log | grep -A9 "K_B" | cut -c1-400
