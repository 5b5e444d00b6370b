set -x
V=$PWD/audio-classification_b200
for lib in libb200sep.so variants/lib_notaps.so variants/lib_noepi.so variants/lib_notaps_noepi.so; do
    B200SEP_LIB=$V/$lib B200SEP_KA=pair timeout 120 python tools/variant_time.py 2>&1 | tail -3 | head -2
done
