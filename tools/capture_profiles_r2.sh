# Round-2 profile capture (run on the GPU box through gpurun, ONE GPU); every ncu pass follows a clean run of the same command.
set -x
export PDM_BENCH_C3=0 PDM_BENCH_C4=0 PDM_BENCH_HBM=0
# 1. launch list of the default bench (headline + screened + 8-bit pixels + denoiser step + C5 slices + e2e)
python bench.py --steps 2 --warmup 3 > gpurun_out/r2_prof_bench.json 2> gpurun_out/r2_prof_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/r2_ncu_launches.log 2>&1
# 2. ncu --set full of single launches at the bench's block shape (168 temperatures = one 6 GiB block of the C2 step)
export PDM_BENCH_NT=168 PDM_BENCH_DENOISER=0 PDM_BENCH_PARITY=0
PDM_BENCH_LATTICE=0 PDM_BENCH_SCREEN=0 python bench.py --steps 1 --warmup 3 > /dev/null 2>&1 || exit 1
PDM_BENCH_LATTICE=0 PDM_BENCH_SCREEN=0 ncu --set full --clock-control none --import-source on -k regex:fused_gemm -s 3 -c 1 \
    -o gpurun_out/prof_r2_fused_f16x3_block python bench.py --steps 1 --warmup 3 > gpurun_out/r2_ncu_full_x3.log 2>&1
PDM_BENCH_LATTICE=0 PDM_BENCH_SCREEN=0 ncu --set full --clock-control none -k regex:noised_rows_philox4 -s 3 -c 1 \
    -o gpurun_out/prof_r2_philox4 python bench.py --steps 1 --warmup 3 > gpurun_out/r2_ncu_full_philox4.log 2>&1
# the screened run: first fused launch of a timed step is the E4M3 stage, then the fp16 one-product stage, then the full pass over a tile list
PDM_BENCH_LATTICE=0 ncu --set full --clock-control none -k regex:fused_gemm_kernelILi2ELi1ELi0ELb0ELb1 -s 2 -c 1 \
    -o gpurun_out/prof_r2_fused_e4m3 python bench.py --steps 1 --warmup 3 > gpurun_out/r2_ncu_full_e4m3.log 2>&1
# 3. the top-k epilogue (k-NN of the dataset against itself, 8192 queries) and the posterior-mean contraction over a tile list
python tools/quick_topk.py > gpurun_out/r2_quick_topk.log 2>&1
ncu --set full --clock-control none -k regex:fused_gemm_kernelILi2ELi3ELi2 -s 1 -c 1 -o gpurun_out/prof_r2_fused_topk \
    python tools/quick_topk.py > gpurun_out/r2_ncu_full_topk.log 2>&1
ls -la gpurun_out/*r2*.ncu-rep
python tools/ncu_summary.py gpurun_out/r2_launches.csv > gpurun_out/r2_launches_summary.csv
for f in fused_f16x3_block philox4 fused_e4m3 fused_topk; do
  python tools/ncu_summary.py gpurun_out/prof_r2_$f.ncu-rep > gpurun_out/r2_${f}_ncu_full.csv 2>&1
done
head -30 gpurun_out/r2_launches_summary.csv
