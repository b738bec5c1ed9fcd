ncu --set full --clock-control none --import-source on -k regex:grouped_linear_lanes --launch-skip 3 --launch-count 1 -o gpurun_out/prof_glin2 python tools/bench_linear.py 32 > gpurun_out/prof_glin2.log 2>&1
tail -2 gpurun_out/prof_glin2.log
