# Round 2, GPU pass J (8 GPUs): sharded check at 8 ranks, bench at N=8 with phase timing, then N=4
set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/check_sharded_gpu.py > gpurun_out/r2j_sharded.log 2>&1; echo "sharded rc=$?"
grep -v "^$" gpurun_out/r2j_sharded.log | grep -v "^\*\|OMP_NUM" | head -14
PDM_BENCH_PHASES=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29558 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2j_bench8.json 2> gpurun_out/r2j_bench8.err; echo "bench8 rc=$?"
grep -E "phase|screened run" gpurun_out/r2j_bench8.err | cut -c1-900
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29554 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2j_bench4.json 2> gpurun_out/r2j_bench4.err; echo "bench4 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2j_bench4.json", "gpurun_out/r2j_bench8.json"):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "headline", j["value"], j["ms_per_step"], "kernel ms", j["roofline"]["kernel_ms_per_step"], "e2e", j["e2e"]["value"], "frac", j["roofline"]["frac"], j["gpu_launches"], j["config"]["sharding"])
        for k in ("parity", "denoiser_step", "c5_trajectory", "c3_hypersphere", "c4_celeba64"):
            print(" ", k, json.dumps(j.get(k))[:1000])
        print("  screened", json.dumps(j["screened"])[:900])
        print("  lattice", j["lattice_8bit"].get("value"), j["lattice_8bit"]["screened"].get("value"), j["lattice_8bit"]["screened"].get("roofline_frac"))
    except Exception as e:
        print(f, "no bench line", e)
PY
