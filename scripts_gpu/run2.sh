cd $GRAFT_REPO_ROOT
timeout 200 python tests/variant_time.py > gpurun_out/vt_fused1.log 2>&1
B200SEP_FUSED=0 timeout 200 python tests/variant_time.py >> gpurun_out/vt_fused1.log 2>&1
cat gpurun_out/vt_fused1.log
timeout 120 python tests/prof_target.py 2 > gpurun_out/plain.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"dw_rs_kernel|apply_in_kernel" -s 60 -c 4 -o gpurun_out/prof_f1 -f python tests/prof_target.py 2 > gpurun_out/ncu_f1.log 2>&1
tail -3 gpurun_out/ncu_f1.log
