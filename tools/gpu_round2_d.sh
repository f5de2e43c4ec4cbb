# Round 2, GPU pass D (8 GPUs): GPU suite with screening on by default, sharded check at 8, bench at N=8 / 4 / 1
set -x
nvidia-smi -L | wc -l
timeout 1800 python -m pytest tests -m gpu -q -rA --durations=8 > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
grep -E "passed|failed|error" gpurun_out/r2d_pytest.log | tail -3
grep -E "^FAILED|^ERROR" gpurun_out/r2d_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/check_sharded_gpu.py > gpurun_out/r2d_sharded.log 2>&1; echo "sharded rc=$?"
grep -v "^$" gpurun_out/r2d_sharded.log | grep -v "^\*\|OMP_NUM" | head -20
for N in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2d_bench$N.json 2> gpurun_out/r2d_bench$N.err; echo "bench$N rc=$?"
tail -c 800 gpurun_out/r2d_bench$N.err
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d_bench1.json 2> gpurun_out/r2d_bench1.err; echo "bench1 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2d_bench1.json", "gpurun_out/r2d_bench4.json", "gpurun_out/r2d_bench8.json"):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "headline", j["value"], j["ms_per_step"], "kernel ms", j["roofline"]["kernel_ms_per_step"], "e2e", j["e2e"]["value"], "frac", j["roofline"]["frac"], j["gpu_launches"], j["config"]["sharding"])
        for k in ("parity", "denoiser_step", "c5_trajectory", "c3_hypersphere", "c4_celeba64"):
            print(" ", k, json.dumps(j.get(k))[:900])
        print("  screened", j["screened"].get("value"), j["screened"].get("roofline_frac"), "lattice", j["lattice_8bit"].get("value"), j["lattice_8bit"]["screened"].get("value"), j["lattice_8bit"]["screened"].get("roofline_frac"))
    except Exception as e:
        print(f, "no bench line", e)
PY
