# compute-sanitizer (memcheck, then racecheck on the small kernels) over the small-shape GPU tests of every kernel family
set -x
export PDM_SAMPLER_GRAPHS=0          # graph capture under the sanitizer is not supported
SEL="tiny_and_ragged or topk_epilogue or delta_rows_shortcut or test_prep_kernels or argmin_ties or empty_queries or test_topk_smallest or test_posterior_mean"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "$SEL" > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitize_memcheck.log | head -20
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_screen.py -m gpu -q -x -k "screened_posterior_mean or ragged" > gpurun_out/sanitize_memcheck_screen.log 2>&1; echo "memcheck screen rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitize_memcheck_screen.log | head -20
