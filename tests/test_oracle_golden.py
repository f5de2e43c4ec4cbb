"""Pin oracle/posterior.py against the fixtures generated from the unmodified
reference (oracle/make_golden.py).  CPU only."""
import math

import torch

from oracle import posterior as orc
from oracle import synthetic as syn
from conftest import load_golden


def close(a, b, rtol=2e-6, atol=1e-6):
    torch.testing.assert_close(a.float(), b.float(), rtol=rtol, atol=atol)


def test_distance_bit_exact():
    g = load_golden("distance.npz")
    assert torch.equal(orc.pairwise_sqdist(g["x"], g["y"]), g["pw_xy"])
    assert torch.equal(orc.pairwise_sqdist(g["x"]), g["pw_xx"])
    assert torch.equal(orc.row_norm_sqr(g["x"].reshape(7, -1)), g["norm_x"])
    assert torch.equal(orc.gram(g["x"].reshape(7, -1), g["y"].reshape(11, -1)), g["gram_xy"])
    d = orc.pairwise_sqdist(g["pts"])
    assert torch.equal(d, g["pw_pts"])
    d.fill_diagonal_(1e10)
    nn1, i1 = d.min(1)
    assert torch.equal(i1, g["nn1_idx"]) and torch.equal(nn1, g["nn1"])
    d.scatter_(1, i1[:, None], 1e10)
    nn2, i2 = d.min(1)
    assert torch.equal(i2, g["nn2_idx"]) and torch.equal(nn2, g["nn2"])


def _check_stats(name):
    g = load_golden(name)
    # RNG order: the oracle's draw reproduces what the reference drew internally
    torch.manual_seed(int(g["seed"]))
    xt = orc.draw_noised_queries(g["x0"], g["temp"], loader_iters="per_temp")
    assert torch.equal(xt, g["xt"])
    torch.manual_seed(int(g["seed"]))
    xm = orc.draw_noised_queries(g["x0"], g["temp"], loader_iters="once_before")
    assert torch.equal(xm, g["xt_metric"])
    ent = orc.entropy_batch(xt, g["data"], g["temp"], chunk=int(g["dl_bs"]))
    assert torch.equal(ent, g["entropy"]), (ent - g["entropy"]).abs().max()
    assert torch.equal(orc.metric_batch(xm, g["data"], g["temp"]), g["metric_plain"])
    assert torch.equal(orc.metric_batch(xm, g["data"], g["temp"], regularize=True), g["metric_global"])
    sig = orc.knn_sigma_reg_sq(g["data"], int(g["knn_k"]), float(g["sigma_reg_scale"]))
    close(orc.metric_batch(xm, g["data"], g["temp"], regularize=True, sigma_reg_sq_per_point=sig),
          g["metric_knn"], rtol=1e-5, atol=1e-7)
    assert abs(orc.dataset_trace_sigma0(g["data"]) - float(g["tr_sigma0"])) <= 1e-6 * float(g["tr_sigma0"])
    return g, xt


def test_stats_gmm():
    _check_stats("stats_gmm.npz")


def test_stats_images():
    _check_stats("stats_images.npz")


def test_stats_clustered():
    _check_stats("stats_clustered.npz")


def test_entropy_forms_agree_and_limits():
    g, xt = _check_stats("stats_gmm.npz")
    n = len(g["data"])
    for i, t in enumerate(g["temp"]):
        energy = 0.5 * orc.pairwise_sqdist(xt[i], g["data"])
        st = orc.boltzmann_rows(energy, t)
        close(st["log_l"] + st["mean_e"] - math.log(n), g["entropy"][i], rtol=1e-4, atol=2e-5)
    # known limits: S -> -log N as T -> 0, S -> 0 as T -> inf
    lo = orc.entropy_batch(g["x0"][None], g["data"], torch.tensor([1e-9]))
    assert torch.allclose(lo, torch.full_like(lo, -math.log(n)), atol=1e-5)
    hi = orc.entropy_batch(g["x0"][None], g["data"], torch.tensor([1e12]))
    assert hi.abs().max() < 1e-4


def test_fp64_loops_cross_check():
    g = load_golden("stats_gmm.npz")
    xt, data = g["xt"][4], g["data"]
    t = g["temp"][4]
    st = orc.boltzmann_rows(0.5 * orc.pairwise_sqdist(xt.double(), data.double()), t.double())
    sl = orc.boltzmann_rows_loops(xt, data, t)
    assert (st["argmin"].numpy() == sl["argmin"]).all()
    for k in ("e_min", "log_l", "mean_e", "mean_e2"):
        torch.testing.assert_close(st[k], torch.from_numpy(sl[k]), rtol=1e-9, atol=1e-9)


def test_denoiser():
    g = load_golden("denoiser.npz")
    for i in range(len(g["taus"])):
        got = orc.posterior_mean_x0(g[f"xt_{i}"], g[f"alpha_bar_{i}"], g["data"])
        assert torch.equal(got, g[f"x0hat_{i}"])
        lt = orc.linear_beta_log_temp(g["taus"][i], float(g["min_temp"]), float(g["max_temp"]))
        close(torch.sigmoid(-lt), g[f"alpha_bar_{i}"], rtol=1e-6, atol=0)
        # fp64 loops: the posterior mean agrees with the direct evaluation
        ab = g[f"alpha_bar_{i}"].double()
        sl = orc.boltzmann_rows_loops(g[f"xt_{i}"].double() / ab.sqrt(), g["data"], (1 - ab) / ab)
        torch.testing.assert_close(got.reshape(10, -1).double(), torch.from_numpy(sl["mean_y"]), rtol=2e-4, atol=2e-5)


def test_metric_utils():
    g = load_golden("metric_utils.npz")
    x, n_y = g["x"], int(g["n_y"])
    torch.manual_seed(int(g["seed"]))
    idx, eps = orc.draw_mc_samples(x, n_y)
    assert torch.equal(idx, g["idx"]) and torch.equal(eps, g["eps"])
    for i in range(3):
        s2 = torch.exp(torch.tensor(float(g[f"scalar_log_sigma_sq_{i}"])))
        y = x[idx] + torch.sqrt(s2) * eps
        close(orc.metric_scalar_from_samples(y, x, s2), g[f"scalar_{i}"], rtol=1e-5, atol=1e-6)
    lam = g["matrix_lambda"]
    evals, evecs = torch.linalg.eigh(torch.diag(lam))
    sigma = evecs @ torch.diag(torch.exp(evals)) @ evecs.t()
    sqrt_sigma = evecs @ torch.diag(torch.sqrt(torch.exp(evals))) @ evecs.t()
    y = x[idx] + (sqrt_sigma @ eps.t()).t()
    close(orc.metric_matrix_from_samples(y, x, torch.diag(sigma)), g["matrix"], rtol=1e-5, atol=1e-6)
    sig = g["rescaled_sigma"]
    y = x[idx] + torch.sqrt(sig) * eps
    close(orc.rescaled_metric_from_samples(y, x, sig), g["rescaled"], rtol=1e-5, atol=1e-6)


def test_cifar_slice():
    g = load_golden("cifar_slice.npz")
    n, b = int(g["n"]), int(g["b"])
    data = syn.uniform_images(n, (3, 32, 32), int(g["data_seed"]))
    assert data.double().sum().item() == float(g["data_checksum"])
    x0 = data[:b].clone()
    torch.manual_seed(int(g["noise_seed"]))
    xt = orc.draw_noised_queries(x0, g["temp"], loader_iters="per_temp")
    assert xt.double().sum().item() == float(g["xt_checksum"])
    assert torch.equal(orc.entropy_batch(xt, data, g["temp"], chunk=512), g["entropy"])
    torch.manual_seed(int(g["noise_seed"]))
    xm = orc.draw_noised_queries(x0, g["temp"], loader_iters="once_before")
    assert torch.equal(orc.metric_batch(xm, data, g["temp"]), g["metric"])
    ab = torch.sigmoid(-orc.linear_beta_log_temp(g["tau"], 1e-4, 2.478e4))
    xq = ab.sqrt() * x0 + (1 - ab).sqrt() * torch.randn(b, 3, 32, 32, generator=syn.gen(int(g["q_seed"])))
    assert xq.double().sum().item() == float(g["xq_checksum"])
    assert torch.equal(orc.posterior_mean_x0(xq, ab, data), g["x0hat"])


def test_oracle_denoiser_gradients_match_reference_autograd():
    """The oracle's ideal denoiser, differentiated by torch autograd, against the gradients the UNMODIFIED reference
    produced for the same inputs (tests/golden/denoiser_grad.npz, oracle/make_golden.py:golden_denoiser_grad)."""
    import math
    g = load_golden("denoiser_grad.npz")
    scale, gamma = 1 + float(g["min_temp"]), math.log((1 + float(g["max_temp"])) / (1 + float(g["min_temp"])))
    for name in [str(c) for c in g["cases"]]:
        data = g[f"{name}_data"]
        for i, tau in enumerate(g["taus"]):
            x = g[f"{name}_xt_{i}"].clone().requires_grad_(True)
            t = tau.view(1).clone().requires_grad_(True)
            log_temp = ((t.pow(2) * gamma).exp() * scale - 1).log()       # diffusion/scheduler/linear.py:11-13
            ab = torch.sigmoid(-log_temp)
            out = orc.posterior_mean_x0(x, ab, data)
            out.backward(g[f"{name}_up_{i}"])
            torch.testing.assert_close(out.detach(), g[f"{name}_x0hat_{i}"], rtol=1e-5, atol=1e-6)
            gs = g[f"{name}_gxt_{i}"].abs().max().item()
            assert (x.grad - g[f"{name}_gxt_{i}"]).abs().max().item() <= 1e-4 * gs + 1e-7, (name, i)
            ts = g[f"{name}_gtau_{i}"].abs().max().item()
            assert (t.grad - g[f"{name}_gtau_{i}"]).abs().max().item() <= 1e-3 * ts + 1e-6, (name, i)
