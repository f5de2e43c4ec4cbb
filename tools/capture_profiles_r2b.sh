# Round-2 profile capture, part 2 (ONE GPU): the kernels whose template arguments select them (ncu matches the mangled name)
set -x
export PDM_BENCH_C3=0 PDM_BENCH_C4=0 PDM_BENCH_HBM=0 PDM_BENCH_NT=168 PDM_BENCH_DENOISER=0 PDM_BENCH_PARITY=0 PDM_BENCH_LATTICE=0
python bench.py --steps 1 --warmup 3 > /dev/null 2>&1 || exit 1
# E4M3 first stage of the screening cascade (second call on: the whole certifiable range in one launch)
ncu --set full --clock-control none --kernel-name-base mangled -k regex:fused_gemm_kernelILi2ELi1ELi0ELb0ELb1 -s 2 -c 1 \
    -o gpurun_out/prof_r2_fused_e4m3 python bench.py --steps 1 --warmup 3 > gpurun_out/r2_ncu_full_e4m3.log 2>&1
# fp16 one-product stage over the device-side tile list the E4M3 stage left
ncu --set full --clock-control none --kernel-name-base mangled -k regex:fused_gemm_kernelILi2ELi1ELi0ELb0ELb0 -s 2 -c 1 \
    -o gpurun_out/prof_r2_fused_f16x1_tiles python bench.py --steps 1 --warmup 3 > gpurun_out/r2_ncu_full_f16x1.log 2>&1
# top-k epilogue
python tools/quick_topk.py > gpurun_out/r2_quick_topk.log 2>&1
ncu --set full --clock-control none --kernel-name-base mangled -k regex:fused_gemm_kernelILi2ELi3ELi2 -s 1 -c 1 -o gpurun_out/prof_r2_fused_topk \
    python tools/quick_topk.py > gpurun_out/r2_ncu_full_topk.log 2>&1
# posterior-mean contraction (EPI_STORE) of a denoiser step
ncu --set full --clock-control none --kernel-name-base mangled -k regex:fused_gemm_kernelILi2ELi3ELi1 -s 3 -c 1 -o gpurun_out/prof_r2_fused_store \
    python tools/quick_denoiser.py > gpurun_out/r2_ncu_full_store.log 2>&1
for f in fused_e4m3 fused_f16x1_tiles fused_topk fused_store; do
  python tools/ncu_summary.py gpurun_out/prof_r2_$f.ncu-rep > gpurun_out/r2_${f}_ncu_full.csv 2>&1
  head -12 gpurun_out/r2_${f}_ncu_full.csv
done
