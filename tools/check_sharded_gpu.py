"""2+ GPU check (torchrun): the row-sharded statistics path with rank-sliced noise generation against the unsharded
engine on the same seed -- same RNG stream, same statistics, same arg-min, same generator position afterwards.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_sharded_gpu.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "physics-of-diffusion-models_b200"))

from pdm_b200 import EmpiricalDataset, PosteriorEngine, EngineConfig  # noqa: E402
from pdm_b200.backend import CudaBackend  # noqa: E402
from pdm_b200.engine import detect_lattice_scale  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    be = CudaBackend(dev)
    n, d, b, n_t = 20_000, 3072, 256, 37                      # 37 temperatures: ragged against any world size
    torch.manual_seed(0)
    data = torch.rand(n, d, device=dev) * 2 - 1
    x0 = data[:b].clone().view(b, 3, 32, 32)
    temps = torch.logspace(-3, 3, n_t, device=dev)
    per = (n + world - 1) // world
    lo, hi = rank * per, min(n, (rank + 1) * per)
    amax = float(data.abs().max().item())
    lat = detect_lattice_scale(be, data, amax)
    full = PosteriorEngine(EmpiricalDataset(data, backend=be), EngineConfig(max_query_bytes=96 << 20))
    gen = PosteriorEngine._cuda_generator(dev)
    ok = True

    def run(engine, seed):
        torch.manual_seed(seed)
        st = engine.noised_stats(x0, temps)
        return st, gen.get_offset()

    ref, off_ref = run(full, 123)
    other = 123 if rank == 0 else 999 + rank                  # with sync_noise every rank must end up on rank 0's stream
    # last entry: certified delta posteriors (EngineConfig.screen) on the sharded dataset
    for slice_noise, sync, seed, screen in ((True, False, 123, False), (False, False, 123, False), (True, True, other, False),
                                            (False, True, other, False), (False, True, other, True)):
        cfg = EngineConfig(max_query_bytes=96 << 20)          # several blocks of temperatures
        cfg.slice_noise, cfg.sync_noise, cfg.screen = slice_noise, sync, screen
        shard = EmpiricalDataset(data[lo:hi], backend=be, index_offset=lo, n_total=n, global_absmax=amax, lattice_scale=lat)
        eng = PosteriorEngine(shard, cfg, group=dist.group.WORLD)
        st, off = run(eng, seed)
        if screen:
            rep = eng.screen_report
            if rank == 0:
                print("screened sharded run:", rep)
            if rep["rows_certified"] < 10 * b or rep["rows_unscreened"] == 0:
                ok = False
                print(f"rank {rank}: screening did not engage as expected: {rep}")
        xn = (x0.reshape(b, -1).double() ** 2).sum(1)[None, :] + d * temps.double()[:, None]
        floor = 8 * 2.0 ** -24 * (xn + (data.double() ** 2).sum(1).max()) / temps.double()[:, None]
        for k in ("log_l", "mean_e", "entropy"):
            err = (st[k].double() - ref[k].double()).abs()
            tol = torch.maximum(2e-5 * ref[k].double().abs() + 2e-6, 0.25 * floor)
            if (err > tol).any():
                ok = False
                print(f"rank {rank} slice={slice_noise} sync={sync}: {k} mismatch, worst {err.max().item():.3e}")
        if not torch.equal(st["argmin"], ref["argmin"]):
            ok = False
            print(f"rank {rank} slice={slice_noise} sync={sync}: argmin mismatch")
        if screen:
            # certified rows recompute E_min in fp64 from the operands: within the fp32 round-off floor of the full pass
            err = (st["e_min"].double() - ref["e_min"].double()).abs()
            if (err > floor * temps.double()[:, None]).any():
                ok = False
                print(f"rank {rank} screened: e_min off by {err.max().item():.3e}")
        elif not torch.equal(st["e_min"], ref["e_min"]):
            ok = False
            print(f"rank {rank} slice={slice_noise} sync={sync}: e_min not bit-identical")
        if not (sync and rank != 0) and off != off_ref:
            ok = False
            print(f"rank {rank} slice={slice_noise} sync={sync}: generator offset {off} != {off_ref}")
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARDED CHECK", "OK" if flag.item() == 1 else "FAILED", f"(world {world})")
    dist.destroy_process_group()
    return 0 if flag.item() == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
