cd $GRAFT_REPO_ROOT
timeout 100 python tests/variant_time.py 2>&1 | tail -1
B200SEP_PDL=0 timeout 100 python tests/variant_time.py 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl.json 2> gpurun_out/bench_pdl.err; echo "bench rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_pdl.json') if l.startswith('{')][0])
print(d['value'], d['ms_per_step'], d['e2e']['value'])"
