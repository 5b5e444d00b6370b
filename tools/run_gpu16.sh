for sp in 1 2 1 2; do
B200SEP_SPLIT=$sp python bench.py --no-sweep --no-eager --no-cpu-baseline 2> gpurun_out/bench16.err > gpurun_out/bench16_sp$sp.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench16_sp$sp.json'))
print('split=$sp', round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], round(d['sustained']['first_steps_ms'],3), round(d['sustained']['lead_in_ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['config']['concurrent_sub_batches'])
PY
done
