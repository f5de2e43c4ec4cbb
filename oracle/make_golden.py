"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Test infrastructure only.  Run in the build container (where /root/reference
exists):   python -m oracle.make_golden
The fixtures pin ``oracle/posterior.py`` (tests/test_oracle_golden.py) and give
the GPU parity tests reference outputs that travel to the GPU box.
Every fixture stores the inputs (or the seeds + a checksum for the larger one),
the noised queries the reference drew internally, and the reference outputs.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
from torch.utils.data import DataLoader, TensorDataset

from oracle import posterior as orc
from oracle import synthetic as syn
from oracle.ref_loader import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _np(d):
    return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in d.items()}


def _save(name, **arrays):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **_np(arrays))
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


def golden_distance(ref):
    g = syn.gen(11)
    x = torch.randn(7, 3, 4, 4, generator=g)
    y = torch.randn(11, 3, 4, 4, generator=g)
    pts = syn.uniform_images(96, (48,), 12)
    d_self = ref.distance.compute_pw_dist_sqr(pts)
    # nearest / second-nearest neighbour flow of scripts/analyze_cifar_nn.py:37-47
    dd = d_self.clone()
    dd.fill_diagonal_(1e10)
    nn1, i1 = dd.min(dim=1)
    dd.scatter_(1, i1.unsqueeze(1), 1e10)
    nn2, i2 = dd.min(dim=1)
    _save("distance.npz", x=x, y=y,
          pw_xy=ref.distance.compute_pw_dist_sqr(x, y),
          pw_xx=ref.distance.compute_pw_dist_sqr(x),
          norm_x=ref.distance.norm_sqr(x.view(7, -1)),
          gram_xy=ref.distance.compute_gram_matrix(x.view(7, -1), y.view(11, -1)),
          pts=pts, pw_pts=d_self, nn1=nn1, nn1_idx=i1, nn2=nn2, nn2_idx=i2)


def _stats_case(ref, name, data, x0, temp, seed, dl_bs, knn_k=5, scale=1.0):
    dl = DataLoader(TensorDataset(data), batch_size=dl_bs, shuffle=False)
    # what the reference will draw internally (incl. its DataLoader passes' CPU-RNG draws)
    torch.manual_seed(seed)
    xt = orc.draw_noised_queries(x0, temp, loader_iters="per_temp")
    torch.manual_seed(seed)
    xt_metric = orc.draw_noised_queries(x0, temp, loader_iters="once_before")
    torch.manual_seed(seed)
    ent = ref.stats.compute_stats_batch(dl, x0, temp)["entropy"]
    out = {"data": data, "x0": x0, "temp": temp, "seed": seed, "dl_bs": dl_bs, "xt": xt,
           "xt_metric": xt_metric, "entropy": ent}
    for tag, kw in (("plain", {}), ("global", {"regularize": True}),
                    ("knn", {"regularize": True, "adaptive_knn": True, "knn_k": knn_k, "sigma_reg_scale": scale})):
        torch.manual_seed(seed)
        out[f"metric_{tag}"] = ref.stats.compute_metric_stats_batch(dl, x0, temp, **kw)["metric_values"]
    out["knn_k"], out["sigma_reg_scale"] = knn_k, scale
    out["tr_sigma0"] = torch.var(data.view(len(data), -1), dim=0).sum()
    _save(name, **out)


def golden_stats(ref):
    d = 16
    data = syn.anisotropic_gmm(d, 3, 300, 7).view(300, 1, d, 1)
    _stats_case(ref, "stats_gmm.npz", data, data[:24].clone(), torch.logspace(-3, 3, 9), 123, 64)
    img = syn.uniform_images(400, (3, 8, 8), 21)
    _stats_case(ref, "stats_images.npz", img, img[100:116].clone(), syn.ddpm_temperatures(1000)[62::125], 321, 100)
    clu = syn.clustered_images(384, (3, 8, 8), 12, 0.02, 22)
    _stats_case(ref, "stats_clustered.npz", clu, clu[:16].clone(), torch.logspace(-4, 2, 10), 77, 128)


def golden_outer_loops(ref):
    """compute_stats / compute_metric_stats incl. the generator loop (utils/stats.py:116-183, 295-311)."""
    data = syn.anisotropic_gmm(8, 2, 120, 3).view(120, 1, 8, 1)
    temp = torch.logspace(-2, 2, 5)
    dl = DataLoader(TensorDataset(data), batch_size=50, shuffle=False)

    def batches():
        i = 0
        while True:
            yield (data[(i * 20) % 120:(i * 20) % 120 + 20],)
            i += 1

    torch.manual_seed(9)
    st = ref.stats.compute_stats(dl, batches(), temp, 60)
    torch.manual_seed(9)
    mt = ref.stats.compute_metric_stats(dl, batches(), temp, 60)
    torch.manual_seed(9)
    mk = ref.stats.compute_metric_stats(dl, batches(), temp, 40, regularize=True, adaptive_knn=True, knn_k=3,
                                        sigma_reg_scale=0.5)
    _save("outer_loops.npz", data=data, temp=temp, seed=9, entropy=st["entropy"], stats_temp=st["temp"],
          metric=mt["metric"], metric_log_temp=mt["log_temp"], metric_tr=mt["dataset_tr_sigma0"],
          metric_knn=mk["metric"])


def golden_denoiser(ref):
    data = syn.uniform_images(256, (3, 8, 8), 31)
    sch = ref.scheduler.LinearBetaScheduler(1e-4, 2.478e4)
    g = syn.gen(32)
    taus = torch.tensor([0.05, 0.3, 0.45, 0.6, 0.95])
    out = {"data": data, "taus": taus, "min_temp": 1e-4, "max_temp": 2.478e4}
    model = ref.true_model.DDPMTrue(sch, "x0", data)
    for i, tau in enumerate(taus):
        ab = sch.alpha_bar_from_tau(tau)
        x0 = data[torch.randint(0, 256, (10,), generator=g)]
        xt = ab.sqrt() * x0 + (1 - ab).sqrt() * torch.randn(10, 3, 8, 8, generator=g)
        out[f"xt_{i}"] = xt
        out[f"alpha_bar_{i}"] = ab
        out[f"x0hat_{i}"] = sch.true_posterior_mean_x0(xt, tau.view(1), data)
        assert torch.equal(out[f"x0hat_{i}"], model(xt, tau.view(1)))
    _save("denoiser.npz", **out)


def golden_denoiser_grad(ref):
    """Autograd of the UNMODIFIED reference through Scheduler.true_posterior_mean_x0 (what
    scripts/optimize_schedule.py:57-91,151 relies on): gradients with respect to xt and to tau."""
    sch = ref.scheduler.LinearBetaScheduler(1e-4, 2.478e4)
    g = syn.gen(51)
    cases = {
        "img": syn.uniform_images(256, (3, 8, 8), 52),                       # d = 192: exact CUDA-core path
        "wide": syn.uniform_images(384, (5, 8, 8), 53),                      # d = 320: tensor path
        "gmm1d": (torch.randn(2000, generator=g) * 0.01 +
                  torch.tensor([-1.1, -0.9, 0.9, 1.1])[torch.randint(0, 4, (2000,), generator=g)]).view(2000, 1, 1, 1),
    }
    taus = torch.tensor([0.08, 0.35, 0.5, 0.7, 0.93])
    out = {"taus": taus, "min_temp": 1e-4, "max_temp": 2.478e4, "cases": np.array(sorted(cases))}
    for name, data in cases.items():
        out[f"{name}_data"] = data
        for i, tau in enumerate(taus):
            ab = sch.alpha_bar_from_tau(tau)
            x0 = data[torch.randint(0, len(data), (12,), generator=g)]
            xt = (ab.sqrt() * x0 + (1 - ab).sqrt() * torch.randn(x0.shape, generator=g)).requires_grad_(True)
            t = tau.view(1).clone().requires_grad_(True)
            up = torch.randn(x0.shape, generator=g)
            x0hat = sch.true_posterior_mean_x0(xt, t, data)
            x0hat.backward(up)
            out[f"{name}_xt_{i}"] = xt.detach()
            out[f"{name}_up_{i}"] = up
            out[f"{name}_x0hat_{i}"] = x0hat.detach()
            out[f"{name}_gxt_{i}"] = xt.grad
            out[f"{name}_gtau_{i}"] = t.grad
    _save("denoiser_grad.npz", **out)


def golden_metric_utils(ref):
    g = syn.gen(41)
    x = torch.randn(200, 3, generator=g) * torch.tensor([1.0, 0.5, 2.0])
    n_y = 400
    torch.manual_seed(5)
    idx, eps = orc.draw_mc_samples(x, n_y)
    out = {"x": x, "n_y": n_y, "seed": 5, "idx": idx, "eps": eps}
    for i, ls in enumerate((-2.0, 0.0, 1.5)):
        torch.manual_seed(5)
        out[f"scalar_{i}"] = ref.metric_utils.compute_metric_scalar(ls, x, n_y)
        out[f"scalar_log_sigma_sq_{i}"] = ls
    lam = torch.tensor([-1.0, 0.0, 0.5])
    torch.manual_seed(5)
    out["matrix"] = ref.metric_utils.compute_metric_matrix(torch.diag(lam), x, n_y)
    out["matrix_lambda"] = lam
    sig = torch.tensor([0.3, 1.0, 2.0])
    torch.manual_seed(5)
    out["rescaled"] = ref.metric_utils.compute_rescaled_metric_matrix(sig, x, n_y)
    out["rescaled_sigma"] = sig
    _save("metric_utils.npz", **out)


def golden_cifar_slice(ref):
    """CIFAR-10-shaped slice (d=3072): inputs are regenerated from seeds at test time."""
    n, b = 2048, 16
    data = syn.uniform_images(n, (3, 32, 32), 51)
    x0 = data[:b].clone()
    temp = syn.ddpm_temperatures(1000)[torch.tensor([9, 199, 399, 599, 799, 999])]
    dl = DataLoader(TensorDataset(data), batch_size=512, shuffle=False)
    torch.manual_seed(52)
    xt = orc.draw_noised_queries(x0, temp, loader_iters="per_temp")
    torch.manual_seed(52)
    ent = ref.stats.compute_stats_batch(dl, x0, temp)["entropy"]
    torch.manual_seed(52)
    met = ref.stats.compute_metric_stats_batch(dl, x0, temp)["metric_values"]
    sch = ref.scheduler.LinearBetaScheduler(1e-4, 2.478e4)
    tau = torch.tensor([0.55])
    ab = sch.alpha_bar_from_tau(tau)
    xq = ab.sqrt() * x0 + (1 - ab).sqrt() * torch.randn(b, 3, 32, 32, generator=syn.gen(53))
    x0hat = sch.true_posterior_mean_x0(xq, tau, data)
    _save("cifar_slice.npz", n=n, b=b, data_seed=51, noise_seed=52, q_seed=53, temp=temp, tau=tau,
          data_checksum=data.double().sum(), xt_checksum=xt.double().sum(), xq_checksum=xq.double().sum(),
          entropy=ent, metric=met, x0hat=x0hat)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    todo = {"distance": golden_distance, "stats": golden_stats, "outer_loops": golden_outer_loops,
            "denoiser": golden_denoiser, "denoiser_grad": golden_denoiser_grad, "metric_utils": golden_metric_utils,
            "cifar_slice": golden_cifar_slice}
    only = sys.argv[1:]                      # python -m oracle.make_golden [name ...]: regenerate a subset
    for name, fn in todo.items():
        if not only or name in only:
            fn(ref)
    return 0


if __name__ == "__main__":
    sys.exit(main())
