cd $GRAFT_REPO_ROOT
timeout 100 python tests/e2e_profile.py 2>&1 | tail -3
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_short.json 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches_v6_split.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
B200SEP_SPLIT=1 timeout 100 python tests/prof_target.py 2 || exit 1
B200SEP_SPLIT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dw_rs_kernel|apply_in_kernel|mask_dec|skip_merge|encoder_kernel|gemm_tc" -s 52 -c 12 -o gpurun_out/prof_f4 -f python tests/prof_target.py 2 > gpurun_out/ncu_f4.log 2>&1
tail -3 gpurun_out/ncu_f4.log
