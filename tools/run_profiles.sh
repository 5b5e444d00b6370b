# Round-2 measurement set (run on a B200 through gpurun): bench lines, launch list, ncu --set full of the block kernels.
set -x
O=gpurun_out
python bench.py > $O/r02f_bench_n1.json 2> $O/r02f_bench_n1.err; tail -c 400 $O/r02f_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02f_bench_reference.json 2> $O/r02f_bench_reference.err
for c in 2 3 4; do python bench.py --config $c --steps 3 --warmup 3 > $O/r02f_bench_config$c.json 2> $O/r02f_bench_config$c.err; tail -c 300 $O/r02f_bench_config$c.err; done
python tools/pit_target.py 5 > $O/r02f_pit_3000.log 2>&1; cat $O/r02f_pit_3000.log
python tools/single_call_latency.py > $O/r02f_single_call_latency.log 2>&1; cat $O/r02f_single_call_latency.log
for ka in single pair; do B200SEP_KA=$ka python tools/variant_time.py; done > $O/r02f_kernel_times.log 2>&1; cat $O/r02f_kernel_times.log
B200SEP_KA=single python tools/power_probe.py 2 3 4 52 > $O/r02f_power_probe.log 2>&1; cat $O/r02f_power_probe.log
CMD="python bench.py --steps 2 --warmup 3 --sustain-s 0 --no-sweep --no-eager --no-cpu-baseline --no-e2e"
$CMD > $O/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 260 --csv --log-file $O/r02f_launches.csv $CMD > $O/ncu1.log 2>&1
B200SEP_SPLIT=1 python tools/prof_target.py 2 > $O/plain2.log 2>&1 && B200SEP_SPLIT=1 ncu --set full --clock-control none --import-source on -k regex:"dw_rs_kernel|apply_in_kernel|mask_dec|skip_merge|encoder" -s 46 -c 10 -o $O/r02f_prof_blocks python tools/prof_target.py 2 > $O/ncu2.log 2>&1
tail -3 $O/ncu1.log $O/ncu2.log
ls -la $O/*.ncu-rep
