set -x
cd $GRAFT_REPO_ROOT
timeout 300 python tests/gpu_diag.py stage16 e2e time > gpurun_out/diag_fused1.log 2>&1
echo "rc=$?" >> gpurun_out/diag_fused1.log
tail -40 gpurun_out/diag_fused1.log
