cd $GRAFT_REPO_ROOT
B200SEP_SPLIT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"apply_in_kernel" -s 26 -c 2 -o gpurun_out/prof_f6_kb -f python tests/prof_target.py 2 > gpurun_out/ncu_f6a.log 2>&1; tail -1 gpurun_out/ncu_f6a.log
B200SEP_SPLIT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mask_dec|skip_merge" -s 2 -c 2 -o gpurun_out/prof_f6_tail -f python tests/prof_target.py 2 > gpurun_out/ncu_f6b.log 2>&1; tail -1 gpurun_out/ncu_f6b.log
