# One-GPU verification pass (run through gpurun): the whole GPU suite, smoke(), the default bench and the reference arm.
#   gpurun --timeout 2400 -- 'bash tools/gpu_verify.sh [tag]'      -> gpurun_out/<tag>_{pytest.log,smoke.log,bench1.json,ref1.json}
tag=${1:-verify}
set -x
timeout 1500 python -m pytest tests -m gpu -q -rA --durations=12 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
grep -E "passed|failed|error" gpurun_out/${tag}_pytest.log | tail -3
grep -E "^FAILED|^ERROR" gpurun_out/${tag}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/${tag}_smoke.log
timeout 900 python bench.py > gpurun_out/${tag}_bench1.json 2> gpurun_out/${tag}_bench1.err; echo "bench1 rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/${tag}_ref1.json 2> gpurun_out/${tag}_ref1.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/${tag}_ref1.json
python tools/bench_digest.py gpurun_out/${tag}_bench1.json
