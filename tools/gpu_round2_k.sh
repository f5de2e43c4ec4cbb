# Round 2, GPU pass K (2 GPUs): whole GPU suite (incl. the 2-GPU grid test), smoke, bench N=2 and N=1 on the final code
set -x
timeout 1800 python -m pytest tests -m gpu -q -rA --durations=8 > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
grep -E "passed|failed|error" gpurun_out/r2k_pytest.log | tail -3
grep -E "^FAILED|^ERROR" gpurun_out/r2k_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2k_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r2k_smoke.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2k_bench2.json 2> gpurun_out/r2k_bench2.err; echo "bench2 rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2k_bench1.json 2> gpurun_out/r2k_bench1.err; echo "bench1 rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2k_ref1.json 2> gpurun_out/r2k_ref1.err; echo "ref rc=$?"; cat gpurun_out/r2k_ref1.json | cut -c1-900
python tools/quick_denoiser.py
