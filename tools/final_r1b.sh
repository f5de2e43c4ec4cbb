set -x
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final.log
tail -3 gpurun_out/pytest_gpu_final.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_final.log
tail -2 gpurun_out/smoke_final.log
timeout 600 python bench.py > gpurun_out/bench_final4.json 2> gpurun_out/bench_final4.err; echo "bench rc=$?"
python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_final4.json"))
print("headline", j["value"], j["ms_per_step"], j["e2e"]["value"], j["roofline"]["frac"], j["gpu_launches"])
s = j["screened"]; print("screened", {k: s[k] for k in s if k != "note"})
s = j["lattice_8bit"]["screened"]; print("lattice", j["lattice_8bit"]["value"], "screened", s["value"], s["ms_per_step"], s.get("roofline_frac"))
PY
