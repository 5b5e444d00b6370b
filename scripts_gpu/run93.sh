cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py > gpurun_out/bench_r1_v9.json 2> gpurun_out/bench_r1_v9.err; tail -1 gpurun_out/bench_r1_v9.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1_v9.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','steps','e2e','gpu_launches','clocks')}); print(d['roofline']['frac'], d['roofline']['traffic'], {k:(round(v['avg_ms']*1e3,1),round(v['frac'],3)) for k,v in d['kernels'].items()})
print(d['cpu_baseline'])
PY
