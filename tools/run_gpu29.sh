timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
B200SEP_KA=single timeout 120 python tools/variant_time.py 2>&1 | tail -3
B200SEP_ENCODER=simt B200SEP_KA=single timeout 120 python tools/variant_time.py 2>&1 | tail -3 | head -1
