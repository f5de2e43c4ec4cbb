# Round 2, GPU pass C (2 GPUs): whole GPU suite, the sharded-grid check
set -x
timeout 1800 python -m pytest tests -m gpu -q -rA --durations=10 > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
grep -E "passed|failed|error" gpurun_out/r2c_pytest.log | tail -5
grep -E "^FAILED|^ERROR" gpurun_out/r2c_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/check_sharded_gpu.py > gpurun_out/r2c_sharded.log 2>&1; echo "sharded rc=$?"
grep -v "^$" gpurun_out/r2c_sharded.log | head -20
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2c_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2c_smoke.log
