"""GPU parity of the reference-facing drop-in modules (utils.distance / utils.stats / utils.metric_utils /
Scheduler.true_posterior_mean_x0 / DDPMTrue) against the reference's golden outputs and the fp64 oracle.
The reference draws its noise on the device RNG; to feed IDENTICAL noised queries the engine's noise hook
replays the CPU draws the golden run used (the noising itself, eps*sqrt(T)+x0, runs in our CUDA kernel and is
bit-identical to torch's)."""
import math

import pytest
import torch
from torch.utils.data import DataLoader, TensorDataset

from conftest import load_golden
from oracle import posterior as orc
from oracle import synthetic as syn
from test_gpu_kernels import arbitrated_close

pytestmark = pytest.mark.gpu


@pytest.fixture()
def replay_noise(cuda_device):
    from pdm_b200 import PosteriorEngine
    import utils.stats as ustats
    ustats._ENGINES.clear()

    def install(eps):
        PosteriorEngine.noise_hook = staticmethod(lambda i, shape, dev: eps[i].reshape(shape).to(dev))
    yield install
    PosteriorEngine.noise_hook = None


def _floor(xt, data, temp):
    xn = (xt.double().reshape(xt.shape[0], xt.shape[1], -1) ** 2).sum(-1)
    yn = (data.double().reshape(len(data), -1) ** 2).sum(1).max()
    return 8 * 2.0 ** -24 * (xn + yn) / temp.double()[:, None]


@pytest.mark.parametrize("name", ["stats_gmm.npz", "stats_images.npz", "stats_clustered.npz"])
def test_compute_stats_batch(replay_noise, name):
    import utils
    g = load_golden(name)
    torch.manual_seed(int(g["seed"]))
    eps = orc.draw_noise(g["x0"].shape, len(g["temp"]), loader_iters="per_temp")
    replay_noise(eps)
    loader = DataLoader(TensorDataset(g["data"]), batch_size=int(g["dl_bs"]), shuffle=False)
    ent = utils.compute_stats_batch(loader, g["x0"], g["temp"])["entropy"]
    assert ent.device.type == "cpu" and ent.shape == g["entropy"].shape
    ref64 = orc.entropy_batch(g["xt"], g["data"], g["temp"], dtype=torch.float64)
    arbitrated_close(ent, g["entropy"], ref64, atol=2e-5, floor=2 * _floor(g["xt"], g["data"], g["temp"]),
                     what=name + " entropy")


@pytest.mark.parametrize("name", ["stats_gmm.npz", "stats_images.npz"])
def test_compute_metric_stats_batch(replay_noise, name, capsys):
    import utils
    g = load_golden(name)
    torch.manual_seed(int(g["seed"]))
    eps = orc.draw_noise(g["x0"].shape, len(g["temp"]), loader_iters="once_before")
    replay_noise(eps)
    loader = DataLoader(TensorDataset(g["data"]), batch_size=int(g["dl_bs"]), shuffle=False)
    sig = orc.knn_sigma_reg_sq(g["data"], int(g["knn_k"]), float(g["sigma_reg_scale"]))
    xm = g["xt_metric"]
    for tag, kw, okw in (("plain", {}, {}), ("global", {"regularize": True}, {"regularize": True}),
                         ("knn", {"regularize": True, "adaptive_knn": True, "knn_k": int(g["knn_k"]),
                                  "sigma_reg_scale": float(g["sigma_reg_scale"])},
                          {"regularize": True, "sigma_reg_sq_per_point": sig})):
        got = utils.compute_metric_stats_batch(loader, g["x0"], g["temp"], **kw)["metric_values"]
        ref64 = orc.metric_batch(xm, g["data"], g["temp"], dtype=torch.float64, **okw)
        st64 = orc.boltzmann_rows(0.5 * orc.pairwise_sqdist(xm[0].double(), g["data"].double()), g["temp"][0].double())
        fl = (_floor(xm, g["data"], g["temp"]) * (1 + 2 * 10.0)).mean(1)
        arbitrated_close(got, g[f"metric_{tag}"], ref64, atol=2e-5, floor=fl, what=f"{name} metric {tag}")
    assert "Tr(Sigma0)=" in capsys.readouterr().out


def test_knn_regulariser_on_gpu(cuda_device):
    from utils.stats import _engine_for, _knn_sigma_reg_sq
    g = load_golden("stats_images.npz")
    loader = DataLoader(TensorDataset(g["data"]), batch_size=100, shuffle=False)
    sig = _knn_sigma_reg_sq(_engine_for(loader), 5, 1.0).cpu()
    torch.testing.assert_close(sig, orc.knn_sigma_reg_sq(g["data"], 5, 1.0), rtol=1e-4, atol=1e-6)


def test_distance_dropin(cuda_device):
    import utils
    g = load_golden("distance.npz")
    for dev in ("cpu", cuda_device):
        x, y = g["x"].to(dev), g["y"].to(dev)
        d = utils.compute_pw_dist_sqr(x, y)
        assert d.device == x.device
        arbitrated_close(d, g["pw_xy"], orc.pairwise_sqdist(g["x"].double(), g["y"].double()), atol=2e-5, what="pw_xy")
    arbitrated_close(utils.compute_pw_dist_sqr(g["x"]), g["pw_xx"], orc.pairwise_sqdist(g["x"].double()), atol=2e-5,
                     what="pw_xx")
    torch.testing.assert_close(utils.norm_sqr(g["x"].reshape(7, -1)), g["norm_x"], rtol=2e-7, atol=0)
    torch.testing.assert_close(utils.compute_gram_matrix(g["x"].reshape(7, -1), g["y"].reshape(11, -1)), g["gram_xy"],
                               rtol=1e-5, atol=1e-5)
    # nearest / second-nearest neighbour flow of scripts/analyze_cifar_nn.py:37-47: indices bit-exact
    d = utils.compute_pw_dist_sqr(g["pts"].to(cuda_device))
    d.fill_diagonal_(1e10)
    nn1, i1 = d.min(dim=1)
    assert torch.equal(i1.cpu(), g["nn1_idx"])
    d.scatter_(1, i1.unsqueeze(1), 1e10)
    assert torch.equal(d.min(dim=1).indices.cpu(), g["nn2_idx"])


def test_denoiser_dropin(cuda_device):
    from diffusion import DDPMTrue
    from diffusion.scheduler import LinearBetaScheduler
    g = load_golden("denoiser.npz")
    sch = LinearBetaScheduler(float(g["min_temp"]), float(g["max_temp"]))
    model = DDPMTrue(sch, "x0", g["data"]).to(cuda_device)
    assert model.train_data.device.type == "cuda"
    for i, tau in enumerate(g["taus"]):
        xt = g[f"xt_{i}"].to(cuda_device)
        got = model(xt, tau.view(1).to(cuda_device))
        assert got.shape == xt.shape and got.device == xt.device and got.dtype == torch.float32
        ref64 = orc.posterior_mean_x0(g[f"xt_{i}"], g[f"alpha_bar_{i}"], g["data"], dtype=torch.float64)
        arbitrated_close(got, g[f"x0hat_{i}"], ref64, atol=2e-5, what=f"x0hat tau#{i}")
        with torch.autocast("cuda", dtype=torch.float16):                  # autocast must not leak in (scheduler.py:58)
            got16 = model(xt.half(), tau.view(1).to(cuda_device))
        assert got16.dtype == torch.float32


def test_denoiser_cifar_slice(cuda_device):
    from diffusion.scheduler import LinearBetaScheduler
    g = load_golden("cifar_slice.npz")
    n, b = int(g["n"]), int(g["b"])
    data = syn.uniform_images(n, (3, 32, 32), int(g["data_seed"]))
    sch = LinearBetaScheduler(1e-4, 2.478e4)
    ab = torch.sigmoid(-orc.linear_beta_log_temp(g["tau"], 1e-4, 2.478e4))
    xq = ab.sqrt() * data[:b] + (1 - ab).sqrt() * torch.randn(b, 3, 32, 32, generator=syn.gen(int(g["q_seed"])))
    got = sch.true_posterior_mean_x0(xq.to(cuda_device), g["tau"].to(cuda_device), data.to(cuda_device))
    ref64 = orc.posterior_mean_x0(xq, ab, data, dtype=torch.float64)
    arbitrated_close(got, g["x0hat"], ref64, atol=5e-5, what="cifar x0hat (tensor path)")


def test_metric_utils_dropin(cuda_device):
    import utils
    g = load_golden("metric_utils.npz")
    x, n_y = g["x"], int(g["n_y"])
    idx, eps = g["idx"], g["eps"]
    # feed the golden run's (idx, eps) draws: patch the two RNG calls of the module
    import utils.metric_utils as mu

    class _Replay:
        def __enter__(self):
            self.ri, self.rn = torch.randint, torch.randn
            mu.torch.randint = lambda *a, **k: idx.to(k.get("device", "cpu"))
            mu.torch.randn = lambda *a, **k: eps.to(k.get("device", "cpu"))

        def __exit__(self, *exc):
            mu.torch.randint, mu.torch.randn = self.ri, self.rn

    with _Replay():
        for i in range(3):
            got = utils.compute_metric_scalar(float(g[f"scalar_log_sigma_sq_{i}"]), x, n_y)
            torch.testing.assert_close(got, g[f"scalar_{i}"], rtol=2e-3, atol=2e-3)
        got = utils.compute_metric_matrix(torch.diag(g["matrix_lambda"]), x, n_y)
        torch.testing.assert_close(got, g["matrix"], rtol=5e-3, atol=5e-3)
        got = utils.compute_rescaled_metric_matrix(g["rescaled_sigma"], x, n_y)
        torch.testing.assert_close(got, g["rescaled"], rtol=5e-3, atol=5e-3)
        xd = x.to(cuda_device)
        got = utils.compute_metric_scalar(0.0, xd, n_y)
        assert got.device.type == "cuda"
        torch.testing.assert_close(got.cpu(), g["scalar_1"], rtol=2e-3, atol=2e-3)


def _grad_close(ours, ref32, ref64, rel, what, floor=0.0):
    """Gradients: within `rel` of the gradient's own scale (max |ref64|) plus an absolute round-off floor, or within
    twice the reference's own fp32 error."""
    ours, ref32, ref64 = ours.detach().double().cpu(), ref32.detach().double().cpu(), ref64.detach().double().cpu()
    tol = torch.maximum(torch.full_like(ref64, rel * (ref64.abs().max().item() + 1e-12) + 1e-7 + floor),
                        2 * (ref32 - ref64).abs())
    err = (ours - ref64).abs()
    assert (err <= tol).all(), f"{what}: worst err {err.max().item():.3e} vs scale {ref64.abs().max().item():.3e}"


@pytest.mark.parametrize("case", ["img", "wide", "gmm1d"])
def test_denoiser_backward_matches_reference_autograd(cuda_device, case):
    """loss.backward() through DDPMTrue / Scheduler.true_posterior_mean_x0 (scripts/optimize_schedule.py:57-91,151):
    engine VJP against the reference's own autograd gradients (golden) with the fp64 oracle as arbiter."""
    import diffusion.scheduler.scheduler as sched
    from diffusion import DDPMTrue
    from diffusion.scheduler import LinearBetaScheduler
    sched._DENOISER_ENGINES.clear()
    g = load_golden("denoiser_grad.npz")
    sch = LinearBetaScheduler(float(g["min_temp"]), float(g["max_temp"]))
    data = g[f"{case}_data"].to(cuda_device)
    model = DDPMTrue(sch, "x0", data)
    for i, tau in enumerate(g["taus"]):
        x = g[f"{case}_xt_{i}"].to(cuda_device).requires_grad_(True)
        t = tau.view(1).to(cuda_device).requires_grad_(True)
        up = g[f"{case}_up_{i}"]
        out = model(x, t)
        assert out.requires_grad and out.shape == x.shape and out.device == x.device
        out.backward(up.to(cuda_device))
        x64 = g[f"{case}_xt_{i}"].double().requires_grad_(True)
        t64 = tau.view(1).double().requires_grad_(True)
        ref = orc.posterior_mean_x0(x64, sch.alpha_bar_from_tau(t64), g[f"{case}_data"], dtype=torch.float64)
        ref.backward(up.double())
        arbitrated_close(out, g[f"{case}_x0hat_{i}"], ref.detach(), atol=2e-5, what=f"{case} x0hat tau#{i}")
        # the weights move by (round-off of an fp32 energy)/T, and the gradient (a covariance under those weights,
        # over T) moves with them: allow that on top of the 2e-4 of the gradient's scale
        ab = sch.alpha_bar_from_tau(tau.double())
        q = g[f"{case}_xt_{i}"].double().reshape(len(up), -1) / ab.sqrt()
        yn = (g[f"{case}_data"].double().reshape(len(data), -1) ** 2).sum(1).max()
        floor_e = (8 * 2.0 ** -24 * ((q ** 2).sum(1).max() + yn) / ((1 - ab) / ab)).item()
        # Cov_p(s, y)/T is a sum of p_j (s_j - a) y_j with s_j = y_j . g evaluated in fp32: near a delta posterior the
        # difference s_j - a cancels, leaving ulp(max |s|) * max |y| / T (times dq/dxt = ab^-1/2) of round-off
        yflat = g[f"{case}_data"].double().reshape(len(data), -1)
        s_max = (up.double().reshape(len(up), -1) @ yflat.t()).abs().max().item()
        floor_g = 8 * 2.0 ** -24 * s_max * yflat.abs().max().item() / ((1 - ab) / ab).item() / ab.sqrt().item()
        _grad_close(x.grad, g[f"{case}_gxt_{i}"], x64.grad, 2e-4 + 2 * floor_e, f"{case} grad xt tau#{i}", floor=floor_g)
        # d/dtau sums g_q . xt over the whole batch (B * d terms carrying the round-off above) times |d ab / d tau|
        tq = tau.view(1).double().requires_grad_(True)
        dab = torch.autograd.grad(sch.alpha_bar_from_tau(tq).sum(), tq)[0].abs().item()
        floor_t = (math.sqrt(up.numel()) * (floor_g / 8) * ab.sqrt().item() * g[f"{case}_xt_{i}"].abs().max().item()
                   * 0.5 * ab.item() ** -1.5 * dab)
        _grad_close(t.grad, g[f"{case}_gtau_{i}"], t64.grad, 1e-3 + 4 * floor_e, f"{case} grad tau tau#{i}", floor=floor_t)
    eng = sched._engine_for_data(data)
    assert eng.precision() == ("exact" if case == "gmm1d" else "f16x3")


def test_denoiser_backward_lattice_and_sampler_chain(cuda_device):
    """Two-product (8-bit image) mode under autograd, through a two-step differentiable DDIM chain like
    optimize_schedule.py's DifferentiableSampler: gradient with respect to the schedule's log-temperatures."""
    import diffusion.scheduler.scheduler as sched
    from diffusion import DDPMTrue, DDPMPredictions
    from diffusion.scheduler import LinearBetaScheduler, cast_log_temp
    sched._DENOISER_ENGINES.clear()
    gen = torch.Generator().manual_seed(5)
    n, shape = 500, (4, 8, 8)
    px = torch.randint(0, 256, (n, *shape), generator=gen, dtype=torch.uint8)
    data = (px.float() / 255 - 0.5) / 0.5
    sch = LinearBetaScheduler(1e-4, 2.478e4)
    x_init = torch.randn(16, *shape, generator=gen)

    def chain(model, log_temp, x0, dtype):
        xt = x0.to(dtype)
        for idx in (1, 0):
            tau = sch.tau_from_log_temp(log_temp[idx]).clip(0, 1)
            ab = cast_log_temp(sch.alpha_bar_from_tau(tau), xt)
            pred = DDPMPredictions(model(xt, tau.view(1)), xt, ab, "x0")
            prev_ab = cast_log_temp(sch.alpha_bar_from_tau(sch.tau_from_log_temp(log_temp[idx - 1]).clip(0, 1)), xt) \
                if idx > 0 else torch.ones_like(ab) * (1 - 1e-6)
            xt = prev_ab.sqrt() * pred.x0 + (1 - prev_ab).sqrt() * pred.eps
        return xt

    model = DDPMTrue(sch, "x0", data.to(cuda_device))
    lt = torch.tensor([-1.0, 1.5], device=cuda_device, requires_grad=True)
    out = chain(model, lt, x_init.to(cuda_device), torch.float32)
    loss = (out ** 2).mean()
    loss.backward()
    assert sched._engine_for_data(model.train_data).precision() == "f16x2"

    class Ref64(torch.nn.Module):
        def forward(self, xt, tau):
            return orc.posterior_mean_x0(xt, sch.alpha_bar_from_tau(tau), data, dtype=torch.float64)

    lt64 = torch.tensor([-1.0, 1.5], dtype=torch.float64, requires_grad=True)
    out64 = chain(Ref64(), lt64, x_init, torch.float64)
    loss64 = (out64 ** 2).mean()
    loss64.backward()
    assert abs(loss.item() - loss64.item()) <= 1e-4 * abs(loss64.item()) + 1e-6
    assert (lt.grad.double().cpu() - lt64.grad).abs().max().item() <= 2e-3 * lt64.grad.abs().max().item() + 1e-7


def test_sharded_grid_two_gpus(cuda_device):
    """Every dataset-shards x temperature-groups layout of the visible GPUs against the unsharded engine, plus both multi-GPU
    samplers against the one-GPU sampler (needs >= 2 GPUs; tools/check_sharded_gpu.py under torchrun)."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "tools", "check_sharded_gpu.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SHARDED CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_c_abi_without_python(cuda_device, tmp_path):
    """The boundary is the C ABI, not Python: a stand-alone C++ host (tests/c/abi_smoke.cu: cudaMalloc'd buffers, plan,
    fused pass on both precisions, merge, error codes) built with nvcc against include/pdm_b200.h + libpdm_b200.so."""
    import os
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available on this box")
    libdir = os.path.join(root, "physics-of-diffusion-models_b200", "lib")
    exe = str(tmp_path / "abi_smoke")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-I", os.path.join(root, "include"),
                    os.path.join(root, "tests", "c", "abi_smoke.cu"), "-o", exe, "-L", libdir, "-lpdm_b200",
                    "-Xlinker", "-rpath", "-Xlinker", libdir], check=True, capture_output=True, text=True, timeout=300)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "ABI SMOKE OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("step_type", ["ddim", "ddpm"])
def test_ideal_sampler_on_gpu(cuda_device, step_type):
    """pdm_b200.IdealSampler (posterior mean + one fused update kernel per step, no host syncs on device values) against
    the reference's DDPMSampler recurrence run through the drop-in DDPMTrue, same CUDA RNG stream."""
    import diffusion.scheduler.scheduler as sched
    from diffusion import DDPMTrue
    from diffusion.scheduler import LinearBetaScheduler
    from pdm_b200 import IdealSampler
    from test_host_logic_cpu import _reference_style_sampling
    sched._DENOISER_ENGINES.clear()
    gen = torch.Generator().manual_seed(3)
    px = torch.randint(0, 256, (800, 3, 16, 16), generator=gen, dtype=torch.uint8)
    data = ((px.float() / 255 - 0.5) / 0.5).to(cuda_device)           # 8-bit images: the two-product path
    sch = LinearBetaScheduler(1e-4, 2.478e4)
    log_temp = sch.log_temp_from_tau(torch.linspace(0, 1, 26, device=cuda_device)[1:])
    sampler = IdealSampler(data, log_temp, step_type=step_type)
    assert sampler.engine.precision() == "f16x2"
    torch.manual_seed(21)
    got = sampler.batch_sample(64)["x"]
    # CUDA graphs: a step configuration met for the second time is captured and replayed from then on; same trajectory
    # as the eager loop (the screening decisions may fall on different steps, both forms are inside the parity tolerance)
    assert sampler.use_graphs and sampler.graph_replays > 0, sampler.graph_replays
    torch.manual_seed(21)
    again = sampler.batch_sample(64)["x"]                               # all replays now
    torch.manual_seed(21)
    eager = IdealSampler(data, log_temp, step_type=step_type, use_graphs=False).batch_sample(64)["x"]
    assert (got - eager).abs().max().item() <= 1e-4 and (again - eager).abs().max().item() <= 1e-4
    model = DDPMTrue(sch, "x0", data)
    torch.manual_seed(21)
    with torch.no_grad():
        want = _reference_style_sampling(model, sch, log_temp, 64, (3, 16, 16), step_type, torch.float32, device=cuda_device)
    assert got.shape == want.shape and got.device == want.device
    err = (got - want).abs().max().item()
    assert err <= 2e-3, err
    # the trajectories end on training images (the posterior is a delta at the lowest noise level)
    d2 = torch.cdist(got.reshape(64, -1), data.reshape(800, -1)).min(1).values
    assert d2.max().item() < 0.5


@pytest.mark.parametrize("step_type,kind", [("ddpm", "continuous"), ("ddim", "pixels")])
def test_ideal_sampler_every_step_against_the_fp64_oracle(cuda_device, step_type, kind):
    """Every state of a fused-sampler trajectory against the ORACLE's fp64 one-step update of the state before it
    (teacher-forced, so the comparison does not depend on how round-off grows along a chaotic trajectory): the ideal
    denoiser of diffusion/scheduler/scheduler.py:58-69 in fp64 on the CPU, the DDPMPredictions algebra and the DDPM / DDIM
    update of diffusion/ddpm_sampling.py:94-110 written out, the noise re-drawn from the same CUDA stream."""
    from diffusion.scheduler import LinearBetaScheduler
    from pdm_b200 import IdealSampler
    gen = torch.Generator().manual_seed(5)
    if kind == "pixels":
        px = torch.randint(0, 256, (900, 3, 8, 8), generator=gen, dtype=torch.uint8)
        data = (px.float() / 255 - 0.5) / 0.5
    else:
        centres = torch.randn(12, 3, 8, 8, generator=gen)
        data = (centres[torch.randint(0, 12, (900,), generator=gen)] + 0.15 * torch.randn(900, 3, 8, 8, generator=gen)).clamp(-1, 1)
    obj, bsz, n_steps = tuple(data.shape[1:]), 48, 40
    sch = LinearBetaScheduler(1e-4, 2.478e4)
    log_temp = sch.log_temp_from_tau(torch.linspace(0, 1, n_steps + 1, dtype=torch.float64)[1:])
    sampler = IdealSampler(data.to(cuda_device), log_temp, step_type=step_type)
    for _ in range(2):                       # second pass: every step configuration replays its CUDA graph
        torch.manual_seed(77)
        states = sampler.batch_sample(bsz, track_states=True)["states"].cpu().double()       # states[idx] = state after step idx
    assert sampler.graph_replays > 0 and states.shape == (n_steps, bsz, *obj)
    torch.manual_seed(77)                    # the draws of that pass, in order: initial state, then one per DDPM step but the last
    x_init = torch.randn(bsz, *obj, device=cuda_device).cpu().double()
    ab_all = torch.sigmoid(-log_temp)
    worst = 0.0
    for idx in range(n_steps - 1, -1, -1):
        xt = x_init if idx == n_steps - 1 else states[idx + 1]
        ab, abp = ab_all[idx], (ab_all[idx - 1] if idx > 0 else torch.tensor(1.0, dtype=torch.float64))
        x0_hat = orc.posterior_mean_x0(xt, ab.reshape(1), data, dtype=torch.float64)
        if step_type == "ddpm":
            alpha = ab / abp
            beta = 1 - alpha
            noise = torch.randn(bsz, *obj, device=cuda_device).cpu().double() if idx > 0 else 0.0
            want = x0_hat * (abp.sqrt() * beta) / (1 - ab) + xt * (alpha.sqrt() * (1 - abp)) / (1 - ab) \
                + noise * ((1 - abp) / (1 - ab) * beta).sqrt()
        else:
            eps = (xt - ab.sqrt() * x0_hat) / (1 - ab).sqrt()
            want = abp.sqrt() * x0_hat + (1 - abp).sqrt() * eps
        # fp32 state: 2^-24 relative round-off of each of the three terms, plus 1e-4 of the denoiser's output scale, or twice
        # what the reference's own fp32 denoiser deviates from fp64 on this state (the arbitration rule of the other tests)
        x0_ref32 = orc.posterior_mean_x0(xt.float(), ab.float().reshape(1), data).double()
        c_x0 = float(abp.sqrt() * (1 - ab / abp) / (1 - ab)) if step_type == "ddpm" else 1.0
        tol = max(1e-4 * max(1.0, float(want.abs().max())), 2 * c_x0 * float((x0_ref32 - x0_hat).abs().max())) \
            + 4 * 2.0 ** -24 * float(xt.abs().max())
        err = float((states[idx] - want).abs().max())
        worst = max(worst, err / tol)
        assert err <= tol, (idx, err, tol)
    print(f"[parity sampler {step_type}/{kind}] worst one-step error / tolerance over {n_steps} steps: {worst:.3f}")
    near = torch.cdist(states[0].reshape(bsz, -1), data.reshape(len(data), -1).double()).min(1).values
    assert near.max().item() < 1e-3                    # the trajectory ends ON training points
