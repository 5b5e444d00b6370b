set -x
B200SEP_KA=pair python bench.py --no-sweep --no-eager --no-cpu-baseline 2> gpurun_out/bench13.err | tee gpurun_out/bench13_pair.json | cut -c1-300
B200SEP_KA=single python bench.py --no-sweep --no-eager --no-cpu-baseline 2> gpurun_out/bench13s.err | tee gpurun_out/bench13_single.json | cut -c1-300
