# Final measurement set of a round (run on a B200 through gpurun): bench lines, launch list, ncu --set full captures.
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r_gpu_tests.log 2>&1
python __graft_entry__.py smoke > gpurun_out/r_smoke.log 2>&1
python bench.py > gpurun_out/r_bench_n1.json 2> gpurun_out/r_bench_n1.err
python bench.py --impl reference > gpurun_out/r_bench_reference.json 2> gpurun_out/r_ref.err
for c in 2 3 4; do python bench.py --config $c > gpurun_out/r_bench_config$c.json 2> gpurun_out/r_c$c.err; done
python tools/pit_target.py > gpurun_out/r_pit_3000.log 2>&1
python tools/variant_time.py > gpurun_out/r_kernel_times.log 2>&1
B200SEP_KA=pair python tools/variant_time.py >> gpurun_out/r_kernel_times.log 2>&1
python tools/single_call_breakdown.py > gpurun_out/r_single_call_breakdown.log 2>&1
python tools/single_call_timeline.py 4.0 > gpurun_out/r_single_call_timeline.log 2>&1
python tools/timeline_tail.py > gpurun_out/r_timeline_mask_dec.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r_launches.csv python bench.py --steps 2 --warmup 3 --sustain-s 0 --no-cpu-baseline --no-e2e --no-sweep --no-eager > gpurun_out/r_ncu_launch.log 2>&1
B200SEP_SPLIT=1 ncu --set full --clock-control none --import-source on -k regex:"encoder_tc|mask_dec|skip_merge|gemm_tc" -c 4 -f -o gpurun_out/r_ends python tools/variant_time.py > gpurun_out/r_ncu_ends.log 2>&1
B200SEP_SPLIT=1 ncu --set full --clock-control none --import-source on -k regex:"apply_in|dw_rs" --launch-skip 40 -c 4 -f -o gpurun_out/r_blocks python tools/variant_time.py > gpurun_out/r_ncu_blocks.log 2>&1
ls -la gpurun_out | tail -20
