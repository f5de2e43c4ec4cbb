# Round-1 (second half) profile capture: every ncu pass follows a clean run of the same command.
set -x
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r1_final2.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches2.log 2>&1
python tools/quick_perf.py --m 172032 --precision f16x1 --iters 2 --configs "2:0:0" > gpurun_out/quick_x1.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:fused_gemm -s 1 -c 1 \
    -o gpurun_out/prof_r1_fused_f16x1_block python tools/quick_perf.py --m 172032 --precision f16x1 --iters 2 --configs "2:0:0" > gpurun_out/ncu_full_x1.log 2>&1
python tools/quick_perf.py --m 172032 --precision f16x3 --iters 2 --configs "2:0:0" > gpurun_out/quick_x3.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:fused_gemm -s 1 -c 1 \
    -o gpurun_out/prof_r1_fused_f16x3_block_b python tools/quick_perf.py --m 172032 --precision f16x3 --iters 2 --configs "2:0:0" > gpurun_out/ncu_full_x3b.log 2>&1
ls -la gpurun_out/*.ncu-rep
