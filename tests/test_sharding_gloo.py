"""gloo runs of the sharded path on the CPU (test double of the backend) over the 2-D grid of pdm_b200/sharding.py:
dataset shards x query groups.  Within a dataset group every rank holds a slice of the dataset rows, reduces its own
partial records, all-gathers one record per row and merges; within a query group the temperatures of the schedule are
dealt round-robin and the results all-gathered.  Whatever the layout, the result must equal the unsharded one on the
same noise stream (the merge is pure math)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

N, D, B = 101, 12, 9


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _inputs():
    g = torch.Generator().manual_seed(0)
    data = torch.randn(N, D, generator=g)
    x0 = data[:B].clone()
    temp = torch.logspace(-2, 2, 7)                       # 7 temperatures: ragged against 2 and 4 query groups
    aux = torch.rand(N, generator=g) + 0.1
    xq = x0 + 0.3 * torch.randn(B, D, generator=torch.Generator().manual_seed(5))
    up = torch.randn(B, D, generator=torch.Generator().manual_seed(6))
    return data, x0, temp, aux, xq, up


def _worker(rank, world, data_shards, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "physics-of-diffusion-models_b200"), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fake_backend import FakeBackend
    from pdm_b200 import EmpiricalDataset, PosteriorEngine, EngineConfig
    from pdm_b200.sharding import make_grid
    data, x0, temp, aux, xq, up = _inputs()
    grid = make_grid(data_shards)
    assert grid.world == world and grid.data_shards == data_shards
    lo, hi = grid.rows(N)
    be = FakeBackend()
    ds = EmpiricalDataset(data[lo:hi], backend=be, index_offset=lo, n_total=N)
    eng = PosteriorEngine(ds, EngineConfig(precision="exact"), group=grid.data_group, query_group=grid.query_group)
    torch.manual_seed(100 + rank)                     # ranks start on DIFFERENT streams: rank 0's state is adopted
    st = eng.noised_stats(x0, temp)
    after = torch.randn(3)                            # the generator is left where the single-process run leaves it
    torch.manual_seed(100 + rank)
    st_aux = eng.noised_stats(x0, temp, aux=aux)      # per-point aux vector over the WHOLE dataset: cut to the shard
    res = {"entropy": st["entropy"], "argmin": st["argmin"], "var_e": st["var_e"], "after": after,
           "aux_mean": st_aux["aux_mean"], "calls": be.calls}
    if grid.query_groups == 1:                        # the denoiser path shards the dataset only
        res["mean"] = eng.posterior_mean(xq, torch.full((B,), 0.5))
        if grid.data_shards > 1:                      # reduce-scatter form: every rank gets its slice of the rows
            part = eng.posterior_mean(xq, torch.full((B,), 0.5), scatter=True)
            per = (B + grid.data_shards - 1) // grid.data_shards
            lo_r = grid.data_index * per
            assert part.shape == (per, D)
            assert torch.allclose(part[:max(0, min(B, lo_r + per) - lo_r)], res["mean"][lo_r:lo_r + per], rtol=1e-6, atol=1e-7)
            assert bool((part[max(0, min(B, lo_r + per) - lo_r):] == 0).all())
        res["gq"], res["gt"] = eng.posterior_mean_backward(xq, torch.full((B,), 0.5), None, up)
    # the reference-facing functions over the same grid: only this rank's rows are uploaded, Tr Sigma_0 / range from
    # all-reduced column moments, k-NN regulariser with the shards taking turns as the query set
    import utils.stats as ustats
    from torch.utils.data import DataLoader, TensorDataset
    os.environ["PDM_SHARD_DATASET"] = "1"
    os.environ["PDM_DATA_SHARDS"] = str(data_shards)
    ustats.default_backend = lambda: be
    ustats._GRID = (world, grid)
    loader = DataLoader(TensorDataset(data.view(N, 1, D, 1)), batch_size=17, shuffle=False)
    e2 = ustats._engine_for(loader)
    assert e2.ds.n == hi - lo and e2.ds.n_total == N and e2.ds.index_offset == lo
    res["sigma_reg_sq"] = ustats._knn_sigma_reg_sq(e2, 5, 1.0)
    res["summary"] = torch.tensor(ustats._dataset_summary(e2))
    torch.manual_seed(100 + rank)
    res["metric_knn"] = ustats.compute_metric_stats_batch(loader, x0.view(B, 1, D, 1), temp, regularize=True, adaptive_knn=True,
                                                          knn_k=5)["metric_values"]
    torch.save(res, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world,data_shards", [(2, 2), (2, 1), (4, 2)])
def test_grid_equals_unsharded(tmp_path, world, data_shards, capsys):
    mp.spawn(_worker, args=(world, data_shards, _free_port(), str(tmp_path)), nprocs=world, join=True)
    ranks = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]
    r0 = ranks[0]
    for r in ranks[1:]:
        for k in r0:
            if k != "calls":
                assert torch.equal(r0[k], r[k]), k            # every rank ends with the same result
    if data_shards > 1:
        assert "reduce" in r0["calls"]

    sys.path.insert(0, HERE)
    from fake_backend import FakeBackend
    from pdm_b200 import EmpiricalDataset, PosteriorEngine, EngineConfig
    from oracle import posterior as orc
    data, x0, temp, aux, xq, up = _inputs()
    eng = PosteriorEngine(EmpiricalDataset(data, backend=FakeBackend()), EngineConfig(precision="exact"))
    torch.manual_seed(100)                                    # rank 0's stream
    st = eng.noised_stats(x0, temp)
    assert torch.equal(torch.randn(3), r0["after"])
    assert torch.equal(st["argmin"], r0["argmin"])
    torch.testing.assert_close(st["entropy"], r0["entropy"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(st["var_e"], r0["var_e"], rtol=1e-4, atol=1e-5)
    torch.manual_seed(100)
    st_aux = eng.noised_stats(x0, temp, aux=aux)
    torch.testing.assert_close(st_aux["aux_mean"], r0["aux_mean"], rtol=1e-5, atol=1e-6)
    assert (st_aux["aux_mean"] - st_aux["aux_mean"].mean()).abs().max() > 1e-3      # the check is not vacuous
    if "mean" in r0:
        torch.testing.assert_close(eng.posterior_mean(xq, torch.full((B,), 0.5)), r0["mean"], rtol=1e-5, atol=1e-6)
        gq, gt = eng.posterior_mean_backward(xq, torch.full((B,), 0.5), None, up)   # backward pass: shards sum to the whole
        torch.testing.assert_close(r0["gq"], gq, rtol=1e-3, atol=1e-4 * gq.abs().max().item())     # fp32 sums, other order
        torch.testing.assert_close(r0["gt"], gt, rtol=1e-3, atol=1e-4 * gt.abs().max().item())
    import utils.stats as ustats
    torch.testing.assert_close(r0["sigma_reg_sq"], ustats._knn_sigma_reg_sq(eng, 5, 1.0), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(r0["sigma_reg_sq"], orc.knn_sigma_reg_sq(data, 5, 1.0), rtol=1e-4, atol=1e-6)
    want = torch.tensor([orc.dataset_trace_sigma0(data), data.min().item(), data.max().item()])
    torch.testing.assert_close(r0["summary"], want, rtol=1e-5, atol=1e-6)
    # the regularised metric through the drop-in on the grid == the oracle on the same noise stream
    torch.manual_seed(100)
    xt = orc.draw_noised_queries(x0, temp)
    ref = orc.metric_batch(xt, data, temp, regularize=True, sigma_reg_sq_per_point=orc.knn_sigma_reg_sq(data, 5, 1.0))
    torch.testing.assert_close(r0["metric_knn"], ref, rtol=2e-3, atol=1e-5)


def _screen_inputs():
    g = torch.Generator().manual_seed(7)
    n, d, b = 240, 64, 16
    data = torch.rand(n, d, generator=g) * 2 - 1
    data[37] = data[5]                                      # a duplicate: the rows of query 5 never certify
    x0 = data[:b].clone()
    temp = torch.logspace(-4, 3, 15)
    noise = torch.randn(len(temp), b, d, generator=g)
    return data, x0, temp, noise


def _screen_worker(rank, world, data_shards, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "physics-of-diffusion-models_b200"), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from test_screen_host_cpu import SplitFakeBackend
    from pdm_b200 import EmpiricalDataset, PosteriorEngine, EngineConfig
    from pdm_b200.sharding import make_grid
    data, x0, temp, noise = _screen_inputs()
    grid = make_grid(data_shards)
    lo, hi = grid.rows(len(data))
    be = SplitFakeBackend()
    ds = EmpiricalDataset(data[lo:hi], backend=be, index_offset=lo, n_total=len(data), global_absmax=float(data.abs().max()))
    cfg = EngineConfig(precision="f16x3", screen=True, screen_f8=True)
    cfg.sync_noise = False
    cfg.max_query_bytes = 4 * x0.shape[0] * data.shape[1] * 12            # blocks of four temperatures
    eng = PosteriorEngine(ds, cfg, group=grid.data_group, query_group=grid.query_group)
    res = {}
    for call in range(4):                 # probing call, then three on the remembered-boundary path (E4M3 mark, dense tiles)
        st = eng.noised_stats(x0, temp, noise_fn=lambda i: noise[i])
        res[f"call{call}"] = {k: st[k] for k in ("entropy", "log_l", "mean_e", "var_e", "e_min", "argmin", "l")}
    res["prior"] = torch.tensor([eng._screen_prior if eng._screen_prior is not None else -1.0,
                                 eng._screen_prior_f8 if eng._screen_prior_f8 is not None else -1.0])
    res["report"] = dict(eng.screen_report)
    res["calls"] = list(be.calls)
    torch.save(res, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world,data_shards", [(2, 2), (2, 1), (4, 2)])
def test_screened_grid_repeated_calls_equal_unsharded(tmp_path, world, data_shards):
    """Certified delta posteriors over the grid, calls one to four on one engine: every rank of a dataset group holds the same
    certificate (merged screening records), so the remembered boundary, the E4M3 stage's mark, the dense tiles of unproven rows
    and their buffer hints agree without any extra exchange; the results equal the unsharded, unscreened engine's."""
    mp.spawn(_screen_worker, args=(world, data_shards, _free_port(), str(tmp_path)), nprocs=world, join=True)
    ranks = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]
    sys.path.insert(0, HERE)
    from test_screen_host_cpu import SplitFakeBackend
    from pdm_b200 import EmpiricalDataset, PosteriorEngine, EngineConfig
    data, x0, temp, noise = _screen_inputs()
    plain = PosteriorEngine(EmpiricalDataset(data, backend=SplitFakeBackend()), EngineConfig(precision="f16x3", screen=False))
    ref = plain.noised_stats(x0, temp, noise_fn=lambda i: noise[i])
    xn = ((noise * temp.sqrt()[:, None, None] + x0[None]) ** 2).sum(-1)
    for r in ranks:
        assert r["prior"][0] > 0, r["prior"]                              # a boundary was found and remembered
        assert r["report"]["rows_certified"] > 0
        assert any(c.startswith("tiles:f16x3:") for c in r["calls"])      # device-side list lengths: the sync-free path ran
        for call in range(4):
            st = r[f"call{call}"]
            for k in ("entropy", "log_l", "mean_e", "var_e"):
                assert torch.isfinite(st[k]).all(), (call, k)
                assert torch.allclose(st[k], ref[k], rtol=1e-4, atol=2e-6), (call, k, (st[k] - ref[k]).abs().max())
            assert torch.equal(st["argmin"], ref["argmin"]), call
            assert ((st["e_min"] - ref["e_min"]).abs() <= 8 * 2.0 ** -24 * (xn + data.shape[1])).all(), call
            assert (st["l"][:4, 5] >= 1.9).all()                           # the duplicate's query is never a certified delta
    for r in ranks[1:]:                                                   # every rank ends with the same numbers
        for call in range(4):
            for k in ("entropy", "argmin", "e_min"):
                assert torch.equal(ranks[0][f"call{call}"][k], r[f"call{call}"][k]), (call, k)
