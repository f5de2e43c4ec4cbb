# Round 2, first GPU pass: the whole GPU suite (incl. the new named-config parity tests), smoke, the bench at N=1
set -x
nvidia-smi -L
nproc; free -g | head -2
timeout 1500 python -m pytest tests -m gpu -q -rA -x --durations=15 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
grep -E "passed|failed|error" gpurun_out/r2a_pytest.log | tail -5
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2a_smoke.log
tail -2 gpurun_out/r2a_smoke.log
PDM_BENCH_WRITE_CURVE=1 timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
cp tests/golden/c2_entropy_curve.npz gpurun_out/ 2>/dev/null
tail -c 1500 gpurun_out/r2a_bench.err
python - <<'PY'
import json
try:
    j = json.load(open("gpurun_out/r2a_bench.json"))
    print("headline", j["value"], j["ms_per_step"], "e2e", j["e2e"]["value"], "frac", j["roofline"]["frac"], j["gpu_launches"])
    for k in ("parity", "denoiser_step", "c5_trajectory", "c3_hypersphere", "c4_celeba64"):
        print(k, json.dumps(j.get(k))[:900])
    print("screened", j["screened"]["value"], j["screened"].get("roofline_frac"))
except Exception as e:
    print("no bench line", e)
PY
