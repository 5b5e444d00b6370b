cd $GRAFT_REPO_ROOT
timeout 300 python bench.py > gpurun_out/bench_r1_v6.json 2> gpurun_out/bench_r1_v6.err; tail -3 gpurun_out/bench_r1_v6.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1_v6.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks','cpu_baseline','roofline')})
PY
