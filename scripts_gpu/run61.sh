cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/bench_r1_v7.json 2> gpurun_out/bench_r1_v7.err; tail -2 gpurun_out/bench_r1_v7.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1_v7.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks')}); print(d['roofline']); print({k:(round(v['avg_ms']*1e3,1),round(v['frac'],3)) for k,v in d['kernels'].items()})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r01_launches_v7.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
B200SEP_SPLIT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dw_rs_kernel|apply_in_kernel|mask_dec|skip_merge|encoder_kernel|gemm_tc" -s 54 -c 10 -o gpurun_out/prof_f5 -f python tests/prof_target.py 2 > gpurun_out/ncu_f5.log 2>&1
tail -2 gpurun_out/ncu_f5.log
