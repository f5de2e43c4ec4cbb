"""2+ GPU check (torchrun): the statistics path over the 2-D grid of pdm_b200/sharding.py (dataset shards x query groups)
against the unsharded engine on the same seed -- same RNG stream, same statistics, same arg-min, bit-identical E_min, same
generator position afterwards -- for every factorisation of the world size, with and without certified delta posteriors, plus
the query-sharded ideal-denoiser sampler against the one-GPU sampler.

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
    full = PosteriorEngine(EmpiricalDataset(data, backend=be), EngineConfig(max_query_bytes=96 << 20, screen=False))
    gen = PosteriorEngine._cuda_generator(dev)
    ok = True

    def run(engine, seed):
        torch.manual_seed(seed)
        st = engine.noised_stats(x0, temps)
        return st, gen.get_offset()

    ref, off_ref = run(full, 123)
    other = 123 if rank == 0 else 999 + rank                  # with sync_noise every rank must end up on rank 0's stream
    from pdm_b200.sharding import make_grid
    shapes = [g for g in range(1, world + 1) if world % g == 0]
    for data_shards in shapes:
        grid = make_grid(data_shards)
        lo, hi = grid.rows(n)
        for sync, seed, screen in ((False, 123, False), (True, other, False), (True, other, True)):
            cfg = EngineConfig(max_query_bytes=96 << 20)          # several blocks of temperatures
            cfg.sync_noise, cfg.screen = sync, screen
            shard = EmpiricalDataset(data[lo:hi], backend=be, index_offset=lo, n_total=n, global_absmax=amax, lattice_scale=lat)
            eng = PosteriorEngine(shard, cfg, group=grid.data_group, query_group=grid.query_group)
            st, off = run(eng, seed)
            tag = f"rank {rank} grid {grid.data_shards}x{grid.query_groups} sync={sync} screen={screen}"
            later = []
            if screen:                 # calls two and three: the remembered-boundary path (no probing, dense tiles, E4M3 mark)
                for call in (2, 3):
                    later.append((call,) + run(eng, seed))
                if eng._screen_prior is None:
                    ok = False
                    print(f"{tag}: no remembered boundary after three calls")
            if screen:
                rep = eng.screen_report
                if rank == 0:
                    print("screened sharded run:", grid.describe(), rep)
                if rep["rows_certified"] < 10 * b // grid.query_groups:
                    ok = False
                    print(f"{tag}: screening did not engage as expected: {rep}")
            xn = (x0.reshape(b, -1).double() ** 2).sum(1)[None, :] + d * temps.double()[:, None]
            floor = 8 * 2.0 ** -24 * (xn + (data.double() ** 2).sum(1).max()) / temps.double()[:, None]
            for k in ("log_l", "mean_e", "entropy"):
                err = (st[k].double() - ref[k].double()).abs()
                tol = torch.maximum(2e-5 * ref[k].double().abs() + 2e-6, 0.25 * floor)
                if (err > tol).any():
                    ok = False
                    print(f"{tag}: {k} mismatch, worst {err.max().item():.3e}")
            if not torch.equal(st["argmin"], ref["argmin"]):
                ok = False
                print(f"{tag}: argmin mismatch")
            if screen:
                # certified rows recompute E_min in fp64 from the operands: within the fp32 round-off floor of the full pass
                err = (st["e_min"].double() - ref["e_min"].double()).abs()
                if (err > floor * temps.double()[:, None]).any():
                    ok = False
                    print(f"{tag}: e_min off by {err.max().item():.3e}")
            elif not torch.equal(st["e_min"], ref["e_min"]):
                ok = False
                print(f"{tag}: e_min not bit-identical")
            if not (sync and rank != 0) and off != off_ref:
                ok = False
                print(f"{tag}: generator offset {off} != {off_ref}")
            for call, st_l, off_l in later:
                for k in ("log_l", "mean_e", "entropy"):
                    err = (st_l[k].double() - ref[k].double()).abs()
                    tol = torch.maximum(2e-5 * ref[k].double().abs() + 2e-6, 0.25 * floor)
                    if (err > tol).any():
                        ok = False
                        print(f"{tag} call {call}: {k} mismatch, worst {err.max().item():.3e}")
                if not torch.equal(st_l["argmin"], ref["argmin"]):
                    ok = False
                    print(f"{tag} call {call}: argmin mismatch")
                err = (st_l["e_min"].double() - ref["e_min"].double()).abs()
                if (err > floor * temps.double()[:, None]).any():
                    ok = False
                    print(f"{tag} call {call}: e_min off by {err.max().item():.3e}")
                if not (sync and rank != 0) and off_l != off_ref:
                    ok = False
                    print(f"{tag} call {call}: generator offset {off_l} != {off_ref}")
            # per-point aux vector over the WHOLE dataset (the k-NN regulariser): every shard must use its own rows
            if not screen:
                aux = torch.rand(n, device=dev, generator=torch.Generator(device=dev).manual_seed(5)) + 0.1
                torch.manual_seed(123)             # the unsharded engine has no group to adopt rank 0's stream from
                a_ref = full.noised_stats(x0, temps, aux=aux)["aux_mean"]
                torch.manual_seed(123 if not sync else seed)
                a_got = eng.noised_stats(x0, temps, aux=aux)["aux_mean"]
                if (a_got - a_ref).abs().max().item() > 1e-4:
                    ok = False
                    print(f"{tag}: aux_mean mismatch {(a_got - a_ref).abs().max().item():.3e}")
    # query-sharded sampler (dataset replicated, trajectories split) == the one-GPU sampler on the same seed
    from pdm_b200 import IdealSampler
    small = data[:4000].view(4000, 3, 32, 32)
    log_temp = torch.linspace(-4.0, 6.0, 12)
    for step_type in ("ddim", "ddpm"):
        torch.manual_seed(55)
        one = IdealSampler(small, log_temp, step_type=step_type).batch_sample(50)["x"]
        torch.manual_seed(55)
        many = IdealSampler(small, log_temp, step_type=step_type, query_group=dist.group.WORLD).batch_sample(50)["x"]
        if not torch.allclose(one, many, rtol=1e-5, atol=1e-5):
            ok = False
            print(f"rank {rank}: query-sharded {step_type} sampler differs by {(one - many).abs().max().item():.3e}")
    # dataset-sharded sampler (shards of the training set, trajectories split over the same group, all-gather of the states
    # + reduce-scatter of the posterior means per step) == the one-GPU sampler
    flat_small = data[:4000]
    per_s = (4000 + world - 1) // world
    s_lo, s_hi = min(4000, rank * per_s), min(4000, (rank + 1) * per_s)
    s_amax = float(flat_small.abs().max().item())
    shard_s = EmpiricalDataset(flat_small[s_lo:s_hi], backend=be, index_offset=s_lo, n_total=4000, global_absmax=s_amax,
                               lattice_scale=detect_lattice_scale(be, flat_small, s_amax))
    eng_s = PosteriorEngine(shard_s, EngineConfig(), group=dist.group.WORLD)
    for step_type in ("ddim", "ddpm"):
        torch.manual_seed(56)
        one = IdealSampler(small, log_temp, step_type=step_type).batch_sample(50)["x"]
        torch.manual_seed(56)
        many = IdealSampler(small, log_temp, step_type=step_type, engine=eng_s).batch_sample(50)["x"]
        if not torch.allclose(one, many, rtol=1e-4, atol=1e-4):
            ok = False
            print(f"rank {rank}: dataset-sharded {step_type} sampler differs by {(one - many).abs().max().item():.3e}")
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARDED CHECK", "OK" if flag.item() == 1 else "FAILED", f"(world {world})")
    dist.destroy_process_group()
    return 0 if flag.item() == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
