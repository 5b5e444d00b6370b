set -x
python tools/pit_target.py 5 > gpurun_out/r02_pit_3000.log 2>&1; cat gpurun_out/r02_pit_3000.log
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 300 gpurun_out/r02_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
CMD="python bench.py --steps 2 --warmup 3 --sustain-s 0 --no-sweep --no-eager --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 260 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu1.log 2>&1
B200SEP_SPLIT=1 python tools/prof_target.py 2 > gpurun_out/plain2.log 2>&1 && B200SEP_SPLIT=1 ncu --set full --clock-control none --import-source on -k regex:"dw_rs_kernel|apply_in_kernel" -s 48 -c 6 -o gpurun_out/r02_prof_blocks python tools/prof_target.py 2 > gpurun_out/ncu2.log 2>&1
python tools/pit_target.py 2 > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pit_moments -s 3 -c 1 -o gpurun_out/r02_prof_pit python tools/pit_target.py 2 > gpurun_out/ncu3.log 2>&1
tail -3 gpurun_out/ncu1.log gpurun_out/ncu2.log gpurun_out/ncu3.log
ls -la gpurun_out/*.ncu-rep
