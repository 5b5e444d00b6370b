cd $GRAFT_REPO_ROOT
timeout 100 python tests/timeline.py 10 > gpurun_out/timeline5.log 2>&1; echo rc=$? >> gpurun_out/timeline5.log
grep -A 30 "CTA 77" gpurun_out/timeline5.log | grep -A3 "K_A" | cut -c1-1500
