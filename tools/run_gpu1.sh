set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for lib in variants/lib_head.so libb200sep.so; do
for ka in pair single; do
    B200SEP_LIB=$PWD/audio-classification_b200/$lib B200SEP_KA=$ka timeout 120 python tools/variant_time.py 2>&1 | tail -3
done; done
B200SEP_KA=pair timeout 120 python tools/timeline.py 6 2>&1 | grep -A 9 "K_A" | cut -c1-1000 | head -24
