# Round 2, GPU pass F (1 GPU): new C3/C4 oracle tests, bench N=1, then the ncu captures
set -x
timeout 1800 python -m pytest tests/test_gpu_named_configs.py tests/test_gpu_screen.py -m gpu -q -rA --durations=8 > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
grep -E "passed|failed|error" gpurun_out/r2f_pytest.log | tail -3
grep -E "^FAILED|^ERROR" gpurun_out/r2f_pytest.log
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r2f_bench1.json 2> gpurun_out/r2f_bench1.err; echo "bench1 rc=$?"
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r2f_bench1.json").read().strip().splitlines()[-1])
print("headline", j["value"], j["ms_per_step"], "e2e", j["e2e"]["value"], "frac", j["roofline"]["frac"])
print("  screened", json.dumps(j["screened"])[:800])
print("  lattice", j["lattice_8bit"].get("value"), j["lattice_8bit"]["screened"].get("value"), j["lattice_8bit"]["screened"].get("roofline_frac"))
print("  c5", json.dumps(j["c5_trajectory"])[:600])
PY
timeout 2400 bash tools/capture_profiles_r2.sh > gpurun_out/r2f_capture.log 2>&1; echo "capture rc=$?"
tail -40 gpurun_out/r2f_capture.log
