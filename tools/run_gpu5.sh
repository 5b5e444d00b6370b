set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -25
for ka in pair single; do
    B200SEP_KA=$ka timeout 120 python tools/variant_time.py 2>&1 | tail -3
done
B200SEP_KA=pair timeout 120 python tools/timeline.py 6 2>&1 | grep -A 9 "K_A" | cut -c1-1200 | head -8
