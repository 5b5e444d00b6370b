cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_fused.json 2> gpurun_out/bench_r1_fused.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_r1_fused.json
