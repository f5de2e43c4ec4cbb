set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded_gpu.py > gpurun_out/check2_screen.log 2>&1; echo "check rc=$?" >> gpurun_out/check2_screen.log
grep -v "^W\|^\[W\|Warning" gpurun_out/check2_screen.log | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > gpurun_out/bench_screen_2gpu.json 2> gpurun_out/bench_screen_2gpu.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_screen_2gpu.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/bench_screen_2gpu.json").read().strip().splitlines()[-1])
print("headline", j["value"], j["ms_per_step"], j["roofline"]["frac"])
print("screened", json.dumps(j.get("screened")))
print("lattice", j["lattice_8bit"]["value"], json.dumps(j["lattice_8bit"].get("screened")))
PY
