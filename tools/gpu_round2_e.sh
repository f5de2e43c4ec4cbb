# Round 2, GPU pass E (2 GPUs): GPU suite, sharded check, bench N=2 and N=1 with phases
set -x
timeout 1800 python -m pytest tests -m gpu -q -rA --durations=8 > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
grep -E "passed|failed|error" gpurun_out/r2e_pytest.log | tail -3
grep -E "^FAILED|^ERROR" gpurun_out/r2e_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/check_sharded_gpu.py > gpurun_out/r2e_sharded.log 2>&1; echo "sharded rc=$?"
grep -v "^$" gpurun_out/r2e_sharded.log | grep -v "^\*\|OMP_NUM" | head -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2e_bench2.json 2> gpurun_out/r2e_bench2.err; echo "bench2 rc=$?"
tail -c 600 gpurun_out/r2e_bench2.err
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/r2e_bench1.json 2> gpurun_out/r2e_bench1.err; echo "bench1 rc=$?"
tail -c 600 gpurun_out/r2e_bench1.err
python - <<'PY'
import json
for f in ("gpurun_out/r2e_bench1.json", "gpurun_out/r2e_bench2.json"):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "headline", j["value"], j["ms_per_step"], "kernel ms", j["roofline"]["kernel_ms_per_step"], "e2e", j["e2e"]["value"], "frac", j["roofline"]["frac"], j["gpu_launches"], j["config"]["sharding"])
        for k in ("denoiser_step", "c5_trajectory"):
            print(" ", k, json.dumps(j.get(k))[:900])
        print("  screened", json.dumps(j["screened"])[:700])
        print("  lattice", j["lattice_8bit"].get("value"), j["lattice_8bit"]["screened"].get("value"), j["lattice_8bit"]["screened"].get("roofline_frac"))
    except Exception as e:
        print(f, "no bench line", e)
PY
