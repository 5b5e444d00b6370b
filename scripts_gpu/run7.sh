cd $GRAFT_REPO_ROOT
timeout 200 python tests/variant_time.py > gpurun_out/vt_fused4.log 2>&1
cat gpurun_out/vt_fused4.log
timeout 300 python tests/gpu_diag.py stage16 e2e 2>&1 | grep -v "fp32\|simt" 
