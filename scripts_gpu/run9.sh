cd $GRAFT_REPO_ROOT
timeout 200 python tests/variant_time.py > gpurun_out/vt_fused5.log 2>&1
cat gpurun_out/vt_fused5.log
timeout 300 python tests/gpu_diag.py e2e 2>&1 | grep "fp16 " 
timeout 200 python tests/timeline.py 10 > gpurun_out/timeline4.log 2>&1; echo rc=$? >> gpurun_out/timeline4.log
grep -A 20 "CTA 77" gpurun_out/timeline4.log | grep -A3 "K_A" | cut -c1-1300
