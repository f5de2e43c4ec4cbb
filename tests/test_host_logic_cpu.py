"""Host logic of the engine and of the reference-facing drop-in modules, on the CPU with the test double of the
backend (tests/fake_backend.py): schedule blocking, RNG order, dataset caching, record merge, posterior mean."""

import os

import pytest
import torch
from torch.utils.data import DataLoader, TensorDataset

from conftest import load_golden
from fake_backend import FakeBackend
from oracle import posterior as orc


@pytest.fixture()
def fake(monkeypatch):
    be = FakeBackend()
    import pdm_b200.engine as engine
    import utils.distance as udist
    import utils.stats as ustats
    import utils.metric_utils as umet
    import diffusion.scheduler.scheduler as sched
    monkeypatch.setattr(engine, "default_backend", lambda: be)
    for mod in (ustats, umet, sched):
        monkeypatch.setattr(mod, "default_backend", lambda: be)
    monkeypatch.setattr(udist, "_backend", be)
    ustats._ENGINES.clear()
    sched._DENOISER_ENGINES.clear()
    return be


def test_engine_blocks_and_merge(fake):
    from pdm_b200 import EmpiricalDataset, PosteriorEngine, EngineConfig
    g = load_golden("stats_gmm.npz")
    data = g["data"].reshape(len(g["data"]), -1)
    ds = EmpiricalDataset(g["data"], backend=fake)
    n_t, b = g["xt"].shape[:2]
    xq = g["xt"].reshape(n_t * b, -1)
    t_rows = g["temp"].repeat_interleave(b)
    ref = orc.boltzmann_rows(0.5 * orc.pairwise_sqdist(xq.double(), data.double()), t_rows.double()[:, None])
    for budget in (1 << 30, 40 * data.shape[1] * 12):          # one block / many small blocks
        eng = PosteriorEngine(ds, EngineConfig(precision="auto", max_query_bytes=budget))
        st = eng.stats(xq, t_rows)
        assert torch.equal(st["argmin"], ref["argmin"])
        torch.testing.assert_close(st["log_l"].double(), ref["log_l"], rtol=2e-3, atol=2e-3)
        torch.testing.assert_close(st["mean_e"].double(), ref["mean_e"], rtol=2e-3, atol=2e-3)
        ent = st["entropy"].view(n_t, b)
        torch.testing.assert_close(ent, g["entropy"], rtol=2e-3, atol=2e-3)


def test_stats_dropin_matches_reference_golden(fake, capsys):
    """compute_stats_batch / compute_metric_stats_batch through the drop-in reproduce the reference's numbers
    (the first call iterates the DataLoader once, then draws one randn per temperature)."""
    import utils
    for name in ("stats_gmm.npz", "stats_images.npz"):
        g = load_golden(name)
        loader = DataLoader(TensorDataset(g["data"]), batch_size=int(g["dl_bs"]), shuffle=False)
        torch.manual_seed(int(g["seed"]))
        ent = utils.compute_stats_batch(loader, g["x0"], g["temp"])["entropy"]
        want = orc.entropy_batch(g["xt_metric"], g["data"], g["temp"])      # same noise stream ("once_before")
        assert ent.shape == want.shape and ent.device.type == "cpu"
        torch.testing.assert_close(ent, want, rtol=2e-3, atol=2e-3)
        # second call re-uses the resident dataset: no further DataLoader pass, so re-seed + one dummy draw
        for tag, kw in (("plain", {}), ("global", {"regularize": True})):
            loader2 = DataLoader(TensorDataset(g["data"]), batch_size=int(g["dl_bs"]), shuffle=False)
            torch.manual_seed(int(g["seed"]))
            got = utils.compute_metric_stats_batch(loader2, g["x0"], g["temp"], **kw)["metric_values"]
            torch.testing.assert_close(got, g[f"metric_{tag}"], rtol=5e-3, atol=1e-5)
        out = capsys.readouterr().out
        assert "Tr(Sigma0)=" in out
        sig = orc.knn_sigma_reg_sq(g["data"], int(g["knn_k"]), float(g["sigma_reg_scale"]))
        loader3 = DataLoader(TensorDataset(g["data"]), batch_size=int(g["dl_bs"]), shuffle=False)
        torch.manual_seed(int(g["seed"]))
        got = utils.compute_metric_stats_batch(loader3, g["x0"], g["temp"], regularize=True, adaptive_knn=True,
                                               knn_k=int(g["knn_k"]), sigma_reg_scale=float(g["sigma_reg_scale"]),
                                               precomputed_sigma_reg_sq=sig)["metric_values"]
        torch.testing.assert_close(got, g["metric_knn"], rtol=5e-3, atol=1e-5)


def test_dataset_is_cached_per_loader(fake):
    import utils
    g = load_golden("stats_gmm.npz")
    loader = DataLoader(TensorDataset(g["data"]), batch_size=64, shuffle=False)
    utils.compute_stats_batch(loader, g["x0"], g["temp"][:2])
    n_norms = fake.calls.count("row_norms")
    utils.compute_stats_batch(loader, g["x0"], g["temp"][:2])
    assert fake.calls.count("row_norms") == n_norms          # no second upload / preparation


def test_outer_loops_and_knn(fake, capsys):
    import utils
    g = load_golden("outer_loops.npz")
    data, temp = g["data"], g["temp"]

    def batches():
        i = 0
        while True:
            yield (data[(i * 20) % 120:(i * 20) % 120 + 20],)
            i += 1

    loader = DataLoader(TensorDataset(data), batch_size=50, shuffle=False)
    torch.manual_seed(int(g["seed"]))
    st = utils.compute_stats(loader, batches(), temp, 60)
    assert set(st) == {"entropy", "temp"} and st["entropy"].shape == temp.shape
    # different noise stream than the reference run (it re-reads the loader per temperature), so compare
    # statistically: entropy means over 60 queries agree within Monte-Carlo error
    assert (st["entropy"] - g["entropy"]).abs().max() < 0.35
    torch.manual_seed(int(g["seed"]))
    mt = utils.compute_metric_stats(loader, batches(), temp, 60)
    assert set(mt) == {"temp", "metric", "log_temp", "dataset_tr_sigma0"}
    torch.testing.assert_close(mt["dataset_tr_sigma0"], g["metric_tr"].float(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(mt["log_temp"], g["metric_log_temp"])
    mk = utils.compute_metric_stats(loader, batches(), temp, 40, regularize=True, adaptive_knn=True, knn_k=3,
                                    sigma_reg_scale=0.5)
    assert mk["metric"].shape == temp.shape and torch.isfinite(mk["metric"]).all()
    # k-NN regulariser against the brute-force oracle
    from utils.stats import _engine_for, _knn_sigma_reg_sq
    sig = _knn_sigma_reg_sq(_engine_for(loader), 3, 0.5)
    torch.testing.assert_close(sig, orc.knn_sigma_reg_sq(data, 3, 0.5), rtol=1e-3, atol=1e-6)


def test_extrapolate_entropy_matches_reference_formula():
    from utils import extrapolate_entropy
    temp = torch.logspace(-3, 2, 12)
    ent = -torch.sigmoid(-(temp.log() + 2)) * 5
    t2, e2 = extrapolate_entropy(temp, ent, 1e-5)
    assert len(t2) == 13 and t2[0].item() == pytest.approx(1e-5)
    lt = t2.log()
    ee = torch.cat([ent[:1], ent])
    slope = (ee[1:] - ee[:-1]) / (lt[1:] - lt[:-1])
    k = int(torch.argmax(slope))
    want = torch.cat(((lt[:k] - lt[k]) * slope[k] + ee[k], ee[k:]))
    torch.testing.assert_close(e2, want)
    t3, e3 = extrapolate_entropy(temp, ent, temp[0].item())
    assert len(t3) == 12


def test_denoiser_dropin(fake):
    from diffusion import DDPMTrue
    from diffusion.scheduler import LinearBetaScheduler
    g = load_golden("denoiser.npz")
    sch = LinearBetaScheduler(float(g["min_temp"]), float(g["max_temp"]))
    model = DDPMTrue(sch, "x0", g["data"])
    assert model.train_data.shape == g["data"].shape
    for i, tau in enumerate(g["taus"]):
        got = model(g[f"xt_{i}"], tau.view(1))
        assert got.shape == g[f"xt_{i}"].shape and got.dtype == torch.float32
        torch.testing.assert_close(got, g[f"x0hat_{i}"], rtol=2e-3, atol=2e-4)
        pred = model.get_predictions(g[f"xt_{i}"], sch.log_temp_from_tau(tau.view(1)))
        torch.testing.assert_close(pred.x0, g[f"x0hat_{i}"], rtol=2e-3, atol=2e-4)


def test_denoiser_backward_chain_rule(fake):
    """Autograd through Scheduler.true_posterior_mean_x0 (scripts/optimize_schedule.py differentiates the sampler
    with respect to the schedule): engine VJP + chain rule against torch autograd on the fp64 oracle."""
    from diffusion.scheduler import LinearBetaScheduler
    g = load_golden("denoiser.npz")
    sch = LinearBetaScheduler(float(g["min_temp"]), float(g["max_temp"]))
    data = g["data"]
    gen = torch.Generator().manual_seed(3)
    for i in (0, len(g["taus"]) // 2, len(g["taus"]) - 1):
        xt = g[f"xt_{i}"][:24]
        up = torch.randn(xt.shape, generator=gen)
        for per_sample in (False, True):
            tau0 = g["taus"][i].double().clamp(1e-3, 1 - 1e-3)
            tau = (tau0.repeat(len(xt)) if per_sample else tau0.view(1)).float().requires_grad_(True)
            x = xt.clone().requires_grad_(True)
            out = sch.true_posterior_mean_x0(x, tau, data)
            out.backward(up)
            # oracle: the reference's formula in fp64 under torch autograd
            x64 = xt.double().requires_grad_(True)
            tau64 = tau.detach().double().requires_grad_(True)
            ab64 = sch.alpha_bar_from_tau(tau64)
            if per_sample:      # the reference's formula takes one noise level per call: evaluate row by row
                ref = torch.cat([orc.posterior_mean_x0(x64[r:r + 1], ab64[r:r + 1], data, dtype=torch.float64)
                                 for r in range(len(xt))])
            else:
                ref = orc.posterior_mean_x0(x64, ab64, data, dtype=torch.float64)
            ref.backward(up.double())
            torch.testing.assert_close(out.detach().double(), ref.detach(), rtol=2e-3, atol=2e-4)
            scale = x64.grad.abs().max().item() + 1e-12
            assert (x.grad.double() - x64.grad).abs().max().item() <= 2e-3 * scale + 1e-6
            tscale = tau64.grad.abs().max().item() + 1e-12
            assert (tau.grad.double() - tau64.grad).abs().max().item() <= 5e-3 * tscale + 1e-6
    # no graph is built when nothing requires grad
    assert not sch.true_posterior_mean_x0(g["xt_0"], g["taus"][:1], data).requires_grad


def test_distance_dropin(fake):
    import utils
    g = load_golden("distance.npz")
    torch.testing.assert_close(utils.compute_pw_dist_sqr(g["x"], g["y"]), g["pw_xy"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(utils.compute_pw_dist_sqr(g["x"]), g["pw_xx"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(utils.norm_sqr(g["x"].reshape(7, -1)), g["norm_x"], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(utils.compute_gram_matrix(g["x"].reshape(7, -1), g["y"].reshape(11, -1)),
                               g["gram_xy"], rtol=1e-5, atol=1e-5)


def test_metric_utils_dropin(fake):
    import utils
    g = load_golden("metric_utils.npz")
    x, n_y = g["x"], int(g["n_y"])
    for i in range(3):
        torch.manual_seed(int(g["seed"]))
        got = utils.compute_metric_scalar(float(g[f"scalar_log_sigma_sq_{i}"]), x, n_y)
        torch.testing.assert_close(got, g[f"scalar_{i}"], rtol=2e-3, atol=2e-3)
    torch.manual_seed(int(g["seed"]))
    got = utils.compute_metric_matrix(torch.diag(g["matrix_lambda"]), x, n_y)
    torch.testing.assert_close(got, g["matrix"], rtol=5e-3, atol=5e-3)
    torch.manual_seed(int(g["seed"]))
    got = utils.compute_rescaled_metric_matrix(g["rescaled_sigma"], x, n_y)
    torch.testing.assert_close(got, g["rescaled"], rtol=5e-3, atol=5e-3)


def test_synthetic_generators():
    import utils
    s = utils.generate_simplex(5)
    assert s.shape == (6, 5)
    d = torch.cdist(s, s)
    off = d[~torch.eye(6, dtype=torch.bool)]
    assert torch.allclose(off, off[0].expand_as(off), atol=1e-5)
    assert utils.generate_cross_polytope(4).shape == (8, 4)
    h = utils.sample_on_hypersphere(16, 100)
    assert torch.allclose(h.norm(dim=1), torch.full((100,), 4.0), atol=1e-4)
    assert utils.generate_dataset("hypersphere", 8).shape == (80, 8)
    with pytest.raises(ValueError):
        utils.generate_dataset("nope")


def test_lattice_candidates_and_detection():
    """Host logic of the lattice (8-bit image) detection: candidate scales and the acceptance rule."""
    from pdm_b200.engine import lattice_candidates, detect_lattice_scale, LATTICE_RATIO_MAX

    c = lattice_candidates(1.0)
    assert c[0] == 2048.0 and c[4:8] == [2040.0, 1020.0, 510.0, 255.0]
    assert all(s * 1.0 <= 2048.0 for s in c)
    assert lattice_candidates(0.0) == [] and lattice_candidates(float("inf")) == []
    assert lattice_candidates(3.0)[0] == 512.0 and lattice_candidates(3.0)[4] == 510.0      # 3 * 510 = 1530 <= 2048

    class Probe:                                                # answers like pdm_lattice_residual_f32 would
        def __init__(self, good):
            self.good, self.asked = good, []

        def lattice_residual(self, y, s):
            self.asked.append(s)
            return torch.tensor([0.0 if s in self.good else 1e-9, 2040.0])

    y = torch.zeros(4, 4)
    p = Probe({510.0, 255.0})
    assert detect_lattice_scale(p, y, absmax=1.0) == 510.0 and p.asked[-3:] == [2040.0, 1020.0, 510.0]
    assert detect_lattice_scale(Probe(set()), y, absmax=1.0) == 0.0
    assert LATTICE_RATIO_MAX == 2.0 ** -44


def _reference_style_sampling(model, sch, log_temp, batch, obj_size, step_type, dtype, device="cpu"):
    """The reference's DDPMSampler.batch_sample / step recurrence (diffusion/ddpm_sampling.py:92-126) written out with
    torch ops, around any DDPM-like ``model`` -- the yardstick for the fused sampler."""
    from diffusion import DDPMPredictions
    from diffusion.scheduler import cast_log_temp, alpha_bar_from_log_temp
    xt = torch.randn(batch, *obj_size, device=device).to(dtype)
    clean = torch.full((1,), -torch.inf, device=device)
    for idx in range(len(log_temp) - 1, -1, -1):
        lt = log_temp[idx].view(1)
        prev = log_temp[idx - 1].view(1) if idx > 0 else clean
        tau = sch.tau_from_log_temp(lt).clip(0, 1)
        ab_pred = cast_log_temp(sch.alpha_bar_from_tau(tau), xt)
        pred = DDPMPredictions(model(xt, tau).to(dtype), xt, ab_pred.to(dtype), "x0")
        ab = cast_log_temp(alpha_bar_from_log_temp(lt), xt).to(dtype)
        abp = cast_log_temp(alpha_bar_from_log_temp(prev), xt).to(dtype)
        if step_type == "ddpm":
            alpha = ab / abp
            beta = 1 - alpha
            noise = torch.randn_like(xt.float()).to(dtype) if idx > 0 else 0
            xt = pred.x0 * (abp.sqrt() * beta) / (1 - ab) + xt * (alpha.sqrt() * (1 - abp)) / (1 - ab) \
                + noise * ((1 - abp) / (1 - ab) * beta).sqrt()
        else:
            xt = abp.sqrt() * pred.x0 + (1 - abp).sqrt() * pred.eps
    return xt


@pytest.mark.parametrize("step_type", ["ddim", "ddpm"])
def test_ideal_sampler_matches_reference_recurrence(fake, step_type):
    """Fused sampler (three coefficients per step, one update kernel) against the reference's step algebra in fp64
    on the same RNG draws."""
    from pdm_b200 import EmpiricalDataset, PosteriorEngine, EngineConfig, IdealSampler, step_coefficients
    from diffusion.scheduler import LinearBetaScheduler
    g = load_golden("denoiser.npz")
    data = g["data"][:96]
    sch = LinearBetaScheduler(float(g["min_temp"]), float(g["max_temp"]))
    log_temp = sch.log_temp_from_tau(torch.linspace(0, 1, 13, dtype=torch.float64)[1:]).float()
    eng = PosteriorEngine(EmpiricalDataset(data, backend=fake), EngineConfig(precision="exact"))
    sampler = IdealSampler(data, log_temp, step_type=step_type, engine=eng)
    torch.manual_seed(11)
    got = sampler.batch_sample(7)["x"]

    class Oracle64(torch.nn.Module):
        def forward(self, xt, tau):
            return orc.posterior_mean_x0(xt, sch.alpha_bar_from_tau(tau.double()), data, dtype=torch.float64)

    torch.manual_seed(11)
    want = _reference_style_sampling(Oracle64(), sch, log_temp.double(), 7, tuple(data.shape[1:]), step_type, torch.float64)
    assert got.shape == want.shape
    torch.testing.assert_close(got.double(), want, rtol=2e-3, atol=2e-3)
    # the last step lands on the posterior mean at the lowest noise level; coefficients at the clean end
    c = step_coefficients(0.3, 1.0, "ddpm")
    assert abs(c[0] - 1.0) < 1e-12 and c[1] == 0.0 and c[2] == 0.0
    out = sampler.sample(10, 4, track_states=True)
    assert out["x"].shape == (10, *data.shape[1:]) and out["states"].shape == (12, 10, *data.shape[1:])


def test_thermo_stats_legacy_schema(fake):
    """compute_thermo_stats: the legacy notebook schema (analyze_stats.ipynb:73-80) out of the same pass --
    S = log_Z + U/T reproduces the entropy of compute_stats, C = var_H/T^2, F = -T log_Z + E_min."""
    import utils
    from torch.utils.data import DataLoader, TensorDataset
    g = load_golden("stats_gmm.npz")
    loader = DataLoader(TensorDataset(g["data"]), batch_size=int(g["dl_bs"]), shuffle=False)

    def batches():
        while True:
            yield (g["x0"],)

    import utils.stats as ustats
    ustats._engine_for(loader)                      # upload once (creating the loader's iterator consumes the global RNG)
    torch.manual_seed(3)
    th = utils.compute_thermo_stats(loader, batches(), g["temp"], n_samples=len(g["x0"]))
    torch.manual_seed(3)
    ent = utils.compute_stats(loader, batches(), g["temp"], n_samples=len(g["x0"]))["entropy"]
    assert set(th) >= {"temp", "entropy", "log_Z", "U", "full_U", "var_H", "heat_capacity", "free_energy"}
    torch.testing.assert_close(th["entropy"], ent, rtol=1e-5, atol=1e-6)
    # the identities hold per query; after averaging over queries they hold for the averages of the same quantities
    torch.testing.assert_close(th["log_Z"] + th["U"] / g["temp"], th["entropy"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(th["var_H"] / g["temp"] ** 2, th["heat_capacity"], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(th["free_energy"], -g["temp"] * th["log_Z"] + (th["full_U"] - th["U"]), rtol=1e-4, atol=1e-4)
    assert (th["var_H"] >= 0).all() and (th["full_U"] >= th["U"] - 1e-6).all()


@pytest.mark.parametrize("parametrization", ["x0", "eps", "score"])
def test_predictions_algebra(parametrization):
    """DDPMPredictions: the three views of one prediction are mutually consistent (xt = sqrt(ab) x0 + sqrt(1-ab) eps,
    score = -eps/sqrt(1-ab)), evaluated on demand, differentiable -- and, where the reference checkout is present, equal to the
    reference's class bit for bit (diffusion/ddpm/ddpm.py:12-30; loaded from its file, no package import needed)."""
    from diffusion import DDPMPredictions
    g = torch.Generator().manual_seed(3)
    xt = torch.randn(5, 3, 4, 4, generator=g)
    pred = torch.randn(5, 3, 4, 4, generator=g, requires_grad=True)
    ab = torch.rand(5, 1, 1, 1, generator=g) * 0.98 + 0.01
    p = DDPMPredictions(pred, xt, ab, parametrization)
    assert p.pred is pred and getattr(p, parametrization) is pred and p.parametrization == parametrization
    assert len(p._cache) == 1                                              # nothing computed until asked for
    torch.testing.assert_close(ab.sqrt() * p.x0 + (1 - ab).sqrt() * p.eps, xt, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(p.score, -p.eps / (1 - ab).sqrt(), rtol=1e-6, atol=1e-6)
    assert p.x0 is p.x0                                                    # cached
    (p.x0.sum() + p.score.sum()).backward()
    assert pred.grad is not None and torch.isfinite(pred.grad).all()
    with pytest.raises(ValueError):
        DDPMPredictions(pred, xt, ab, "v")
    ref_file = "/root/reference/diffusion/ddpm/ddpm.py"
    if os.path.exists(ref_file):
        src = open(ref_file).read()
        body = src[src.index("class DDPMPredictions"):src.index("class DDPM(")]
        ns = {"Tensor": torch.Tensor}
        exec(body, ns)                                                     # the reference's class, verbatim, stand-alone
        r = ns["DDPMPredictions"](pred.detach(), xt, ab, parametrization)
        for k in ("x0", "eps", "score"):
            assert torch.equal(getattr(p, k).detach(), getattr(r, k)), k
