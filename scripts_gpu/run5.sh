cd $GRAFT_REPO_ROOT
timeout 200 python tests/timeline.py 10 > gpurun_out/timeline.log 2>&1; echo rc=$? >> gpurun_out/timeline.log
head -70 gpurun_out/timeline.log
