cd $GRAFT_REPO_ROOT
timeout 200 python tests/variant_time.py > gpurun_out/vt_fused3.log 2>&1
cat gpurun_out/vt_fused3.log
timeout 300 python tests/gpu_diag.py e2e 2>&1 | grep "fp16 " 
timeout 200 python tests/timeline.py 18 > gpurun_out/timeline2.log 2>&1; echo rc=$? >> gpurun_out/timeline2.log
grep -A9 "CTA 77" gpurun_out/timeline2.log | cut -c1-900
