cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py > gpurun_out/bench_r1_split.json 2> gpurun_out/bench_r1_split.err; tail -3 gpurun_out/bench_r1_split.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1_split.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks','cpu_baseline')})
PY
timeout 100 python tests/e2e_breakdown.py 2>&1 | tail -3
