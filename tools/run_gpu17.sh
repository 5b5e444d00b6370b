timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_range.py -m gpu -x -q 2>&1 | tail -8
B200SEP_KA=single timeout 120 python tools/variant_time.py 2>&1 | tail -3
B200SEP_SPLIT=1 B200SEP_KA=single timeout 120 python tools/variant_time.py 2>&1 | tail -3 | head -1
