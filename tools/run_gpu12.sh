set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for ka in single pair; do B200SEP_KA=$ka timeout 120 python tools/variant_time.py 2>&1 | tail -3; done
python bench.py --no-sweep 2> gpurun_out/bench12.err | tee gpurun_out/bench12.json | cut -c1-600
