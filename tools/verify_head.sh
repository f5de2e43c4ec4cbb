# HEAD verification on the GPU box: GPU parity tests, smoke(), HBM-kernel probe, ncu captures of the norm / merge
# kernels (each after a clean run of the same command), default bench.
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_head.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_head.log
tail -3 gpurun_out/pytest_gpu_head.log
python __graft_entry__.py smoke > gpurun_out/smoke_head.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_head.log
tail -2 gpurun_out/smoke_head.log
python tools/quick_hbm.py > gpurun_out/hbm_head.log 2>&1; tail -20 gpurun_out/hbm_head.log
ncu --set full --clock-control none --import-source on -k regex:^row_norms_kernel -s 1 -c 1 \
    -o gpurun_out/prof_r1_row_norms python tools/quick_hbm.py > gpurun_out/ncu_full_norms.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:merge_partials_kernel -s 1 -c 1 \
    -o gpurun_out/prof_r1_merge python tools/quick_hbm.py > gpurun_out/ncu_full_merge.log 2>&1
python bench.py > gpurun_out/bench_head.json 2> gpurun_out/bench_head.err; echo "bench rc=$?"
cat gpurun_out/bench_head.json | head -c 1500
ls -la gpurun_out/*.ncu-rep
