for ka in pair single; do
    B200SEP_KA=$ka timeout 120 python tools/variant_time.py 2>&1 | tail -3
done
