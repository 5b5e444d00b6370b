set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -25
B200SEP_KA=single timeout 120 python tools/variant_time.py 2>&1 | tail -3
