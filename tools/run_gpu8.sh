set -x
V=$PWD/audio-classification_b200
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -25
for lib in libb200sep.so variants/lib_s4.so; do
    B200SEP_LIB=$V/$lib B200SEP_KA=pair timeout 120 python tools/variant_time.py 2>&1 | tail -3 | head -2
done
B200SEP_KA=pair timeout 120 python tools/timeline.py 6 2>&1 | grep -A 9 "K_A" | cut -c1-1200 | head -8
