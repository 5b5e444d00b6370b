cd $GRAFT_REPO_ROOT
timeout 300 python tests/gpu_diag.py stage16 e2e > gpurun_out/diag_fused2.log 2>&1
echo "rc=$?" >> gpurun_out/diag_fused2.log
cat gpurun_out/diag_fused2.log | grep -v "fp32\|simt"
timeout 200 python tests/variant_time.py > gpurun_out/vt_fused2.log 2>&1
cat gpurun_out/vt_fused2.log
