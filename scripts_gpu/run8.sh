cd $GRAFT_REPO_ROOT
timeout 200 python tests/timeline.py 10 > gpurun_out/timeline3.log 2>&1; echo rc=$? >> gpurun_out/timeline3.log
grep -A 20 "CTA 77" gpurun_out/timeline3.log | grep -A4 "K_A" | cut -c1-1300
