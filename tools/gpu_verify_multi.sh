# Multi-GPU verification pass (gpurun --gpus N): every factorisation of the grid against the unsharded engine, both
# multi-GPU samplers (tools/check_sharded_gpu.py), then the bench at N ranks exactly as the driver launches it.
#   gpurun --gpus 8 --timeout 2400 -- 'bash tools/gpu_verify_multi.sh 8 [tag]'
n=${1:-2}; tag=${2:-verify}
set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 tools/check_sharded_gpu.py > gpurun_out/${tag}_sharded${n}.log 2>&1; echo "sharded rc=$?"
grep -v "^$" gpurun_out/${tag}_sharded${n}.log | grep -v "^\*\|OMP_NUM" | tail -16
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29558 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${tag}_bench${n}.json 2> gpurun_out/${tag}_bench${n}.err; echo "bench$n rc=$?"
python tools/bench_digest.py gpurun_out/${tag}_bench${n}.json
