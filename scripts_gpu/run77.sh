cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/bench_r1_v8.json 2> gpurun_out/bench_r1_v8.err; tail -2 gpurun_out/bench_r1_v8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1_v8.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks')}); print({k:(round(v['avg_ms']*1e3,1),round(v['frac'],3)) for k,v in d['kernels'].items()})
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | cut -c1-200
