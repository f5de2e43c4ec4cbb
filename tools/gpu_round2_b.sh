# Round 2, GPU pass B (2 GPUs): whole GPU suite (no -x), the sharded-grid check, bench at N=2 and N=1
set -x
nvidia-smi -L
timeout 1800 python -m pytest tests -m gpu -q -rA --durations=15 > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
grep -E "passed|failed|error" gpurun_out/r2b_pytest.log | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/check_sharded_gpu.py > gpurun_out/r2b_sharded.log 2>&1; echo "sharded rc=$?"
tail -5 gpurun_out/r2b_sharded.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2b_bench2.json 2> gpurun_out/r2b_bench2.err; echo "bench2 rc=$?"
tail -c 1200 gpurun_out/r2b_bench2.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench1.json 2> gpurun_out/r2b_bench1.err; echo "bench1 rc=$?"
tail -c 600 gpurun_out/r2b_bench1.err
python - <<'PY'
import json
for f in ("gpurun_out/r2b_bench1.json", "gpurun_out/r2b_bench2.json"):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "headline", j["value"], j["ms_per_step"], "e2e", j["e2e"]["value"], "frac", j["roofline"]["frac"], j["gpu_launches"], j["config"]["sharding"])
        for k in ("parity", "denoiser_step", "c5_trajectory", "c3_hypersphere", "c4_celeba64"):
            print(" ", k, json.dumps(j.get(k))[:700])
        print("  screened", j["screened"].get("value"), j["screened"].get("roofline_frac"))
    except Exception as e:
        print(f, "no bench line", e)
PY
