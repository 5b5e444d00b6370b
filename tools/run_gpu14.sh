set -x
timeout 100 python tools/pit_target.py 5
for kc in 0 1; do
B200SEP_KB_CLUSTER=$kc B200SEP_KA=pair timeout 120 python tools/variant_time.py 2>&1 | tail -3 | head -1
B200SEP_KB_CLUSTER=$kc B200SEP_KA=pair python bench.py --no-sweep --no-eager --no-cpu-baseline --no-e2e 2> gpurun_out/bench14.err > gpurun_out/bench14_kc$kc.json
done
