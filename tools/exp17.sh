set -x
timeout 400 python -m pytest tests/test_gpu_screen.py -m gpu -x -q 2>&1 | tail -15
PDM_BENCH_DENOISER=0 PDM_BENCH_HBM=0 timeout 300 python bench.py > gpurun_out/bench_f8.json 2> gpurun_out/bench_f8.err
tail -3 gpurun_out/bench_f8.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_f8.json"))
print("headline", j["value"], j["ms_per_step"])
s = j["screened"]; print("screened", s["value"], s["ms_per_step"], s["rows_screened_per_step"], s["rows_certified_per_step"], s["row_tiles_full_pass_per_step"], s["gpu_launches"])
s = j["lattice_8bit"]["screened"]; print("lattice", j["lattice_8bit"]["value"], "screened", s["value"], s["ms_per_step"], s.get("roofline_frac"))
PY
