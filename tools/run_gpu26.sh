V=$PWD/audio-classification_b200
for lib in libb200sep.so variants/lib_hint2000.so variants/lib_hint200.so; do
    B200SEP_LIB=$V/$lib B200SEP_KA=single timeout 120 python tools/variant_time.py 2>&1 | tail -3
done
