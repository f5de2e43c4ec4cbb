"""World-size-2 gloo run of the sharded path on the CPU (test double of the backend): every rank holds half of
the dataset rows, sees all queries, reduces its own partial records, all-gathers one record per row and
merges.  The result must equal the unsharded one (the merge is pure math)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "physics-of-diffusion-models_b200"), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fake_backend import FakeBackend
    from pdm_b200 import EmpiricalDataset, PosteriorEngine, EngineConfig
    g = torch.Generator().manual_seed(0)
    n, d, b = 101, 12, 9
    data = torch.randn(n, d, generator=g)
    x0 = data[:b].clone()
    temp = torch.logspace(-2, 2, 5)
    per = (n + world - 1) // world
    lo, hi = rank * per, min(n, (rank + 1) * per)
    be = FakeBackend()
    ds = EmpiricalDataset(data[lo:hi], backend=be, index_offset=lo, n_total=n)
    eng = PosteriorEngine(ds, EngineConfig(precision="exact"), group=dist.group.WORLD)
    torch.manual_seed(100 + rank)                     # ranks draw DIFFERENT noise: rank 0's is broadcast
    st = eng.noised_stats(x0, temp)
    xq = x0 + 0.3 * torch.randn(b, d, generator=torch.Generator().manual_seed(5))
    mean = eng.posterior_mean(xq, torch.full((b,), 0.5))
    up = torch.randn(b, d, generator=torch.Generator().manual_seed(6))
    gq, gt = eng.posterior_mean_backward(xq, torch.full((b,), 0.5), None, up)
    # adaptive k-NN regulariser on the sharded dataset: every rank searches its shard, candidates are merged
    import utils.stats as ustats
    ds.full_moments_source = data
    sig = ustats._knn_sigma_reg_sq(eng, 5, 1.0)
    torch.save({"entropy": st["entropy"], "argmin": st["argmin"], "var_e": st["var_e"], "mean": mean, "sigma_reg_sq": sig,
                "gq": gq, "gt": gt,
                "calls": be.calls}, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_sharded_equals_unsharded(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    for k in ("entropy", "argmin", "var_e", "mean", "sigma_reg_sq", "gq", "gt"):
        assert torch.equal(r0[k], r1[k]), k                  # both ranks end with the same merged result
    assert "reduce" in r0["calls"]

    sys.path.insert(0, HERE)
    from fake_backend import FakeBackend
    from pdm_b200 import EmpiricalDataset, PosteriorEngine, EngineConfig
    g = torch.Generator().manual_seed(0)
    n, d, b = 101, 12, 9
    data = torch.randn(n, d, generator=g)
    x0 = data[:b].clone()
    temp = torch.logspace(-2, 2, 5)
    eng = PosteriorEngine(EmpiricalDataset(data, backend=FakeBackend()), EngineConfig(precision="exact"))
    torch.manual_seed(100)                                    # rank 0's stream
    st = eng.noised_stats(x0, temp)
    assert torch.equal(st["argmin"], r0["argmin"])
    torch.testing.assert_close(st["entropy"], r0["entropy"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(st["var_e"], r0["var_e"], rtol=1e-4, atol=1e-5)
    xq = x0 + 0.3 * torch.randn(b, d, generator=torch.Generator().manual_seed(5))
    torch.testing.assert_close(eng.posterior_mean(xq, torch.full((b,), 0.5)), r0["mean"], rtol=1e-5, atol=1e-6)
    up = torch.randn(b, d, generator=torch.Generator().manual_seed(6))
    gq, gt = eng.posterior_mean_backward(xq, torch.full((b,), 0.5), None, up)     # backward pass: shards sum to the whole
    torch.testing.assert_close(r0["gq"], gq, rtol=1e-3, atol=1e-4 * gq.abs().max().item())     # fp32 sums, other order
    torch.testing.assert_close(r0["gt"], gt, rtol=1e-3, atol=1e-4 * gt.abs().max().item())
    from oracle import posterior as orc
    import utils.stats as ustats
    torch.testing.assert_close(r0["sigma_reg_sq"], ustats._knn_sigma_reg_sq(eng, 5, 1.0), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(r0["sigma_reg_sq"], orc.knn_sigma_reg_sq(data, 5, 1.0), rtol=1e-4, atol=1e-6)
