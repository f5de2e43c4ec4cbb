# Round-2 profile capture, part 3 (ONE GPU), on the final code: launch list of the bench and the reworked weights kernel.
# Every ncu pass follows a clean run of the same command.
set -x
export PDM_BENCH_C3=0 PDM_BENCH_C4=0 PDM_BENCH_HBM=0 PDM_BENCH_C5_FULL=0
python bench.py --steps 2 --warmup 3 > gpurun_out/r2c_prof_bench.json 2> gpurun_out/r2c_prof_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2c_launches.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/r2c_ncu_launches.log 2>&1
python tools/ncu_summary.py gpurun_out/r2c_launches.csv > gpurun_out/r2c_launches_summary.csv
head -24 gpurun_out/r2c_launches_summary.csv
# weights_kernel<false> of a C5 step (10 000 x 50 000 energies -> fp16 hi/lo weights)
python tools/quick_denoiser.py > gpurun_out/r2c_quick_denoiser.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:weights_kernel -s 2 -c 1 -o gpurun_out/prof_r2c_weights \
    python tools/quick_denoiser.py > gpurun_out/r2c_ncu_full_weights.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_r2c_weights.ncu-rep > gpurun_out/r2c_weights_ncu_full.csv 2>&1
cat gpurun_out/r2c_weights_ncu_full.csv | head -30
