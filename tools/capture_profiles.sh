# Round-1 profile capture (run on the GPU box through gpurun); every ncu pass follows a clean run of the same command.
set -x
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1_final.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
export PDM_BENCH_NT=168 PDM_BENCH_DENOISER=0        # 168 temperatures = one 6 GiB block of the C2 step
PDM_BENCH_LATTICE=0 ncu --set full --clock-control none --import-source on -k regex:fused_gemm -s 2 -c 1 \
    -o gpurun_out/prof_r1_fused_f16x3_block python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_full_x3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_gemm -s 5 -c 1 \
    -o gpurun_out/prof_r1_fused_f16x2_block python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_full_x2.log 2>&1
ncu --set full --clock-control none -k regex:noised_rows_philox -s 2 -c 1 \
    -o gpurun_out/prof_r1_philox python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_full_philox.log 2>&1
ls -la gpurun_out/*.ncu-rep
