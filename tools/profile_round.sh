set -x
python bench.py --impl reference > gpurun_out/g_bench_reference.json 2> gpurun_out/g_ref.err
for c in 2 3 4; do python bench.py --config $c > gpurun_out/g_bench_config$c.json 2> gpurun_out/g_c$c.err; done
python tools/pit_target.py > gpurun_out/g_pit_3000.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/g_launches.csv python bench.py --steps 2 --warmup 3 --sustain-s 0 --no-cpu-baseline --no-e2e --no-sweep --no-eager > gpurun_out/g_ncu_launch.log 2>&1
B200SEP_SPLIT=1 ncu --set full --clock-control none --import-source on -k regex:encoder_tc -c 1 -f -o gpurun_out/g_enc python tools/variant_time.py > gpurun_out/g_ncu_enc.log 2>&1
ls -la gpurun_out | tail -12
