set -x
V=$PWD/audio-classification_b200
for lib in libb200sep.so variants/lib_nopf.so variants/lib_nohalo.so variants/lib_raw6.so; do
    B200SEP_LIB=$V/$lib B200SEP_KA=pair timeout 120 python tools/variant_time.py 2>&1 | tail -3
done
B200SEP_KA=pair timeout 120 python tools/timeline.py 6 2>&1 | grep -A 9 "K_A" | cut -c1-1200 | head -8
B200SEP_LIB=$V/variants/lib_raw6.so B200SEP_KA=pair timeout 120 python tools/timeline.py 6 2>&1 | grep -A 9 "K_A" | cut -c1-1200 | head -8
