set -x
timeout 1800 python -m pytest tests -m gpu -q -rA --durations=8 > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
grep -E "passed|failed|error" gpurun_out/r2i_pytest.log | tail -3
grep -E "^FAILED|^ERROR" gpurun_out/r2i_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2i_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r2i_smoke.log
python tools/quick_denoiser.py
timeout 1500 bash tools/capture_profiles_r2b.sh > gpurun_out/r2i_capture.log 2>&1; echo "capture rc=$?"
tail -n 60 gpurun_out/r2i_capture.log
