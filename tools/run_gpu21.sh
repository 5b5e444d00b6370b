B200SEP_SPLIT=1 python tools/prof_target.py 2 > gpurun_out/plain21.log 2>&1 && B200SEP_SPLIT=1 ncu --set full --clock-control none --import-source on -k regex:"mask_dec_kernel|apply_in_kernel|encoder" -s 50 -c 4 -o gpurun_out/r02b_prof_tail python tools/prof_target.py 2 > gpurun_out/ncu21.log 2>&1
tail -3 gpurun_out/ncu21.log; ls -la gpurun_out/*.ncu-rep
