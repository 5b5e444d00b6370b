cd $GRAFT_REPO_ROOT
timeout 120 python tests/prof_target.py 2 > gpurun_out/plain.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"dw_rs_kernel|apply_in_kernel|mask_dec|skip_merge" -s 62 -c 6 -o gpurun_out/prof_f3 -f python tests/prof_target.py 2 > gpurun_out/ncu_f3.log 2>&1
tail -3 gpurun_out/ncu_f3.log
