# First GPU run of the certified-delta-posterior path: its tests, then the whole GPU suite, then the bench.
set -x
timeout 600 python -m pytest tests/test_gpu_screen.py -m gpu -x -q > gpurun_out/pytest_screen.log 2>&1; echo "screen rc=$?" >> gpurun_out/pytest_screen.log
tail -40 gpurun_out/pytest_screen.log
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_head.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_head.log
tail -5 gpurun_out/pytest_gpu_head.log
timeout 600 python bench.py > gpurun_out/bench_screen.json 2> gpurun_out/bench_screen.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_screen.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_screen.json"))
print("headline", j["value"], j["ms_per_step"], j["roofline"]["frac"])
print("screened", json.dumps(j.get("screened")))
print("lattice", j["lattice_8bit"]["value"], json.dumps(j["lattice_8bit"].get("screened")))
PY
