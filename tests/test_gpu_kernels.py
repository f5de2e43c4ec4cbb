"""GPU parity tests of the CUDA kernels behind the C ABI, against the CPU oracle (fp32 restatement of the
reference, with its fp64 re-evaluation as arbiter).  Run on the B200 box:  pytest -m gpu

Tolerance contract (BASELINE.json north_star): min-shifted log l, <e>, <e^2>, Var,
entropy, posterior mean and E_min within 1e-4 relative of the reference on identical inputs, where the
fp64 oracle arbitrates when the reference's own fp32 cancellation noise is larger:
    |ours - ref64| <= max(rtol*|ref64| + atol, 2*|ref32 - ref64|, floor);  arg-min indices bit-exact.
`floor` is the a-priori round-off of ANY fp32 evaluation of ||x||^2 - 2x.y + ||y||^2 (the reference's
formula, utils/distance.py:21): 8 * 2^-24 * (||x||^2 + ||y||^2) on an energy, divided by T on the
min-shifted exponents.  The reference's own outputs move by that much between BLAS builds, so no
implementation can be held closer to the fp64 value than that.
"""
import math

import pytest
import torch

from oracle import posterior as orc
from oracle import synthetic as syn
from conftest import load_golden

pytestmark = pytest.mark.gpu

RTOL = 1e-4


@pytest.fixture(scope="module")
def backend(cuda_device):
    from pdm_b200.backend import CudaBackend
    return CudaBackend(cuda_device)


def arbitrated_close(ours, ref32, ref64, rtol=RTOL, atol=1e-6, what="", floor=None):
    ours = ours.detach().double().cpu()
    ref32 = ref32.detach().double().cpu()
    ref64 = ref64.detach().double().cpu()
    tol = torch.maximum(rtol * ref64.abs() + atol, 2 * (ref32 - ref64).abs())
    if floor is not None:
        tol = torch.maximum(tol, floor.double().reshape(tol.shape) if floor.numel() == tol.numel() else floor.double())
    err = (ours - ref64).abs()
    bad = err > tol
    assert not bad.any(), (f"{what}: {int(bad.sum())} of {bad.numel()} outside tolerance; worst err "
                           f"{err[bad].max().item():.3e} vs tol {tol[bad][err[bad].argmax()].item():.3e}")


def oracle_rows(xq, data, temp_rows, aux=None):
    """fp32 and fp64 oracle statistics for explicit query rows with per-row temperatures."""
    res = {}
    for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
        e = 0.5 * orc.pairwise_sqdist(xq.to(dt), data.to(dt))
        st = orc.boltzmann_rows(e, temp_rows.to(dt)[:, None], aux=None if aux is None else aux.to(dt))
        st["entropy"] = st["log_l"] + st["mean_e"] - math.log(len(data))
        res[name] = st
    # fp32 round-off floor of the norm expansion: 8 * 2^-24 * (||x||^2 + max ||y||^2) on E, / T on e
    xn = (xq.double().reshape(len(xq), -1) ** 2).sum(1)
    yn = (data.double().reshape(len(data), -1) ** 2).sum(1).max()
    res["floor_E"] = 8 * 2.0 ** -24 * (xn + yn)
    res["floor_e"] = res["floor_E"] / temp_rows.double()
    return res


def run_stats(backend, xq, data, temp_rows, precision, aux=None, cta_group=0, n_splits=0, m_group=0):
    dev = backend.device
    y = data.reshape(len(data), -1).float().to(dev).contiguous()
    x = xq.reshape(len(xq), -1).float().to(dev).contiguous()
    inv_t = (1.0 / temp_rows.float()).to(dev).contiguous()
    y_norm = backend.row_norms(y)
    auxd = None if aux is None else aux.float().to(dev).contiguous()
    if precision == "exact":
        prep = backend.prepare_rows(x, len(x), want_x=True, want_split=False)
        parts = backend.posterior_stats(precision="exact", M=len(x), N=len(y), d=y.shape[1], q_norm=prep["norms"],
                                        y_norm=y_norm, inv_temp=inv_t, q=prep["x"], y=y, y_aux=auxd, n_splits=n_splits)
    else:
        from pdm_b200.engine import pow2_scale_for
        scale = pow2_scale_for(float(backend.absmax(y).item()))
        ys = backend.prepare_rows(y, len(y), fixed_scale=scale, want_norms=False)
        prep = backend.prepare_rows(x, len(x))
        parts = backend.posterior_stats(precision=precision, M=len(x), N=len(y), d=y.shape[1], q_norm=prep["norms"],
                                        y_norm=y_norm, inv_temp=inv_t, q_split=(prep["hi"], prep["lo"], prep["inv_scale"]),
                                        y_split=(ys["hi"], ys["lo"]), y_inv_scale=1.0 / scale, y_aux=auxd,
                                        cta_group=cta_group, n_splits=n_splits, m_group=m_group)
    out, argmin = backend.merge(parts, inv_t, len(y))
    torch.cuda.synchronize()
    return out.cpu(), argmin.cpu()


def check_stats(out, argmin, ref, aux=False, what=""):
    from pdm_b200 import _cabi as cabi
    r32, r64 = ref["f32"], ref["f64"]
    # bit-exact arg-min wherever the reference's own fp32 and fp64 evaluations agree on it
    agree = r32["argmin"] == r64["argmin"]
    assert torch.equal(argmin[agree], r64["argmin"][agree]), what + ": argmin mismatch"
    fE, fe = ref["floor_E"], ref["floor_e"]
    fe2 = fe * (1 + 2 * r64["mean_e"].double())
    arbitrated_close(out[cabi.OUT_E_MIN], r32["e_min"], r64["e_min"], atol=1e-5, what=what + " e_min", floor=fE)
    arbitrated_close(out[cabi.OUT_LOG_L], r32["log_l"], r64["log_l"], atol=1e-5, what=what + " log_l", floor=fe)
    arbitrated_close(out[cabi.OUT_MEAN_E], r32["mean_e"], r64["mean_e"], atol=1e-5, what=what + " mean_e", floor=fe)
    arbitrated_close(out[cabi.OUT_MEAN_E2], r32["mean_e2"], r64["mean_e2"], atol=1e-5, what=what + " mean_e2", floor=fe2)
    arbitrated_close(out[cabi.OUT_VAR_E], r32["var_e"], r64["var_e"], atol=2e-5, what=what + " var_e", floor=fe2)
    arbitrated_close(out[cabi.OUT_ENTROPY], r32["entropy"], r64["entropy"], atol=2e-5, what=what + " entropy", floor=2 * fe)
    if aux:
        arbitrated_close(out[cabi.OUT_AUX_MEAN], r32["aux_mean"], r64["aux_mean"], atol=1e-7, what=what + " aux")


# ------------------------------------------------------------------------------------------------
def test_prep_kernels(backend):
    dev = backend.device
    g = syn.gen(1)
    for rows, d in ((37, 50), (64, 192), (5, 3072)):
        x = torch.randn(rows, d, generator=g)
        xd = x.to(dev)
        ref = (x.double() ** 2).sum(1)
        got = backend.row_norms(xd).cpu().double()
        assert ((got - ref).abs() <= 1.2e-7 * ref).all()
        noise = torch.randn(3 * rows, d, generator=g)
        sigma = torch.rand(3 * rows, generator=g) + 0.1
        post = torch.rand(3 * rows, generator=g) + 0.5
        r = backend.prepare_rows(xd, 3 * rows, noise=noise.to(dev), sigma=sigma.to(dev), post=post.to(dev), want_x=True)
        want = (noise * sigma[:, None] + x.repeat(3, 1)) * post[:, None]
        assert torch.equal(r["x"].cpu(), want), "noising must match torch's mul-then-add bit for bit"
        nref = (want.double() ** 2).sum(1)
        assert ((r["norms"].cpu().double() - nref).abs() <= 1.2e-7 * nref).all()
        # hi + lo reconstructs v * 2^k to ~22 bits of the row maximum
        inv = r["inv_scale"].cpu().double()[:, None]
        rec = (r["hi"].cpu().double()[:, :d] + r["lo"].cpu().double()[:, :d]) * inv
        amax = want.abs().max(1).values.double()[:, None]
        assert ((rec - want.double()).abs() <= amax * 2.0 ** -21).all()
        scaled_max = (want.abs().max(1).values.double() / inv[:, 0])
        assert ((scaled_max >= 2048) & (scaled_max < 4096)).all()
        assert (r["hi"].cpu()[:, d:] == 0).all() and (r["lo"].cpu()[:, d:] == 0).all()
    y = torch.randn(70, 45, generator=g)
    hi, lo = backend.transpose_split(y.to(dev), 64.0)
    rec = (hi.cpu().double() + lo.cpu().double())[:, :70] / 64.0
    assert (rec - y.t().double()).abs().max() < 2.0 ** -20 * y.abs().max()
    assert (hi.cpu()[:, 70:] == 0).all()
    s, s2, mm = backend.column_moments(y.to(dev))
    torch.testing.assert_close(s.cpu(), y.double().sum(0), rtol=1e-12, atol=1e-12)
    torch.testing.assert_close(s2.cpu(), (y.double() ** 2).sum(0), rtol=1e-12, atol=1e-12)
    assert mm[0].item() == y.min().item() and mm[1].item() == y.max().item()
    assert backend.absmax(y.to(dev)).item() == y.abs().max().item()


@pytest.mark.parametrize("name", ["stats_gmm.npz", "stats_images.npz", "stats_clustered.npz"])
def test_exact_stats_golden(backend, name):
    g = load_golden(name)
    data = g["data"].reshape(len(g["data"]), -1)
    n_t, b = g["xt"].shape[:2]
    xq = g["xt"].reshape(n_t * b, -1)
    t_rows = g["temp"].repeat_interleave(b)
    sig = orc.knn_sigma_reg_sq(g["data"], int(g["knn_k"]), float(g["sigma_reg_scale"]))
    ref = oracle_rows(xq, data, t_rows, aux=sig)
    for splits in (1, 3):
        out, argmin = run_stats(backend, xq, data, t_rows, "exact", aux=sig, n_splits=splits)
        check_stats(out, argmin, ref, aux=True, what=f"{name} exact S={splits}")
    # the reference's own batch outputs (entropy per query, utils/stats.py:289)
    from pdm_b200 import _cabi as cabi
    ent = out[cabi.OUT_ENTROPY].view(n_t, b)
    ref64 = ref["f64"]["entropy"].view(n_t, b)
    arbitrated_close(ent, g["entropy"], ref64, atol=2e-5, what=name + " entropy vs reference golden",
                     floor=(2 * ref["floor_e"]).view(n_t, b))


@pytest.mark.parametrize("cta_group", [1, 2])
@pytest.mark.parametrize("precision", ["f16x3"])
def test_tensor_stats_small(backend, cta_group, precision):
    """Tensor path at shapes that exercise ragged tiles: M, N, d not multiples of the tile sizes."""
    g = syn.gen(5)
    for (m, n, d) in ((200, 700, 192), (300, 1000, 328), (130, 260, 64)):
        data = torch.rand(n, d, generator=g) * 2 - 1
        x0 = data[torch.randint(0, n, (m,), generator=g)]
        t_rows = torch.logspace(-3, 3, m)
        xq = x0 + t_rows.sqrt()[:, None] * torch.randn(m, d, generator=g)
        aux = torch.rand(n, generator=g) * 0.1
        ref = oracle_rows(xq, data, t_rows, aux=aux)
        for splits, grp in ((0, 0), (1, 1), (3, 2)):
            out, argmin = run_stats(backend, xq, data, t_rows, precision, aux=aux, cta_group=cta_group,
                                    n_splits=splits, m_group=grp)
            check_stats(out, argmin, ref, aux=True, what=f"tensor cg={cta_group} ({m},{n},{d}) S={splits}")


@pytest.mark.parametrize("cta_group", [1, 2])
def test_tensor_stats_cifar_slice(backend, cta_group):
    g = load_golden("cifar_slice.npz")
    n, b = int(g["n"]), int(g["b"])
    data = syn.uniform_images(n, (3, 32, 32), int(g["data_seed"]))
    x0 = data[:b].clone()
    torch.manual_seed(int(g["noise_seed"]))
    xt = orc.draw_noised_queries(x0, g["temp"], loader_iters="per_temp")
    n_t = len(g["temp"])
    xq = xt.reshape(n_t * b, -1)
    t_rows = g["temp"].repeat_interleave(b)
    ref = oracle_rows(xq, data.reshape(n, -1), t_rows)
    out, argmin = run_stats(backend, xq, data, t_rows, "f16x3", cta_group=cta_group)
    check_stats(out, argmin, ref, what=f"cifar slice cg={cta_group}")
    from pdm_b200 import _cabi as cabi
    arbitrated_close(out[cabi.OUT_ENTROPY].view(n_t, b), g["entropy"], ref["f64"]["entropy"].view(n_t, b), atol=2e-5,
                     what="cifar slice entropy vs reference golden", floor=(2 * ref["floor_e"]).view(n_t, b))


def test_tensor_accumulation_error(backend):
    """How far the tensor-core Gram entries are from fp64 (documents the accumulation behaviour)."""
    g = syn.gen(9)
    n, d, m = 512, 3072, 128
    data = torch.rand(n, d, generator=g) * 2 - 1
    xq = data[:m] + 0.05 * torch.randn(m, d, generator=g)
    dev = backend.device
    from pdm_b200.engine import pow2_scale_for
    y = data.to(dev)
    scale = pow2_scale_for(float(backend.absmax(y).item()))
    ys = backend.prepare_rows(y, n, fixed_scale=scale, want_norms=False)
    prep = backend.prepare_rows(xq.to(dev), m)
    y_norm = backend.row_norms(y)
    for prec in ("f16x3", "f16x1"):
        dist = torch.empty(m, n, device=dev)
        backend.posterior_stats(precision=prec, M=m, N=n, d=d, q_norm=prep["norms"], y_norm=y_norm, inv_temp=None,
                                q_split=(prep["hi"], prep["lo"], prep["inv_scale"]), y_split=(ys["hi"], ys["lo"]),
                                y_inv_scale=1.0 / scale, want_partials=False, energy_out=dist, energy_mult=2.0)
        ref64 = orc.pairwise_sqdist(xq.double(), data.double())
        ref32 = orc.pairwise_sqdist(xq, data)
        err = (dist.cpu().double() - ref64)
        e32 = (ref32.double() - ref64)
        print(f"\n[{prec}] dist err vs fp64: max {err.abs().max():.3e} mean {err.mean():.3e} rms {err.pow(2).mean().sqrt():.3e}"
              f" | reference fp32: max {e32.abs().max():.3e} mean {e32.mean():.3e} rms {e32.pow(2).mean().sqrt():.3e}")
        if prec == "f16x3":
            assert err.abs().max() < 20 * max(e32.abs().max().item(), 1e-4)


def test_posterior_mean(backend):
    from pdm_b200 import EmpiricalDataset, PosteriorEngine, EngineConfig
    g = load_golden("denoiser.npz")
    ds = EmpiricalDataset(g["data"], backend=backend)
    for prec in ("exact", "f16x3"):
        eng = PosteriorEngine(ds, EngineConfig(precision=prec))
        for i in range(len(g["taus"])):
            ab = g[f"alpha_bar_{i}"].double()
            xt = g[f"xt_{i}"]
            t_rows = ((1 - ab) / ab).float().expand(len(xt))
            post = (1 / ab.sqrt()).float().expand(len(xt))
            got = eng.posterior_mean(xt, t_rows, post=post).cpu()
            ref64 = orc.posterior_mean_x0(xt, g[f"alpha_bar_{i}"], g["data"], dtype=torch.float64).reshape(len(xt), -1)
            arbitrated_close(got, g[f"x0hat_{i}"].reshape(len(xt), -1), ref64, atol=2e-5, what=f"x0hat {prec} tau#{i}")


def test_pairwise_dense(backend):
    from pdm_b200 import EmpiricalDataset, PosteriorEngine, EngineConfig
    g = load_golden("distance.npz")
    ds = EmpiricalDataset(g["pts"], backend=backend)
    eng = PosteriorEngine(ds, EngineConfig(precision="exact"))
    d = eng.pairwise_sqdist(g["pts"]).cpu()
    ref64 = orc.pairwise_sqdist(g["pts"].double())
    arbitrated_close(d, g["pw_pts"], ref64, atol=2e-5, what="dense self distance")
    d.fill_diagonal_(1e10)
    nn1, i1 = d.min(1)
    assert torch.equal(i1, g["nn1_idx"])
    d.scatter_(1, i1[:, None], 1e10)
    assert torch.equal(d.min(1).indices, g["nn2_idx"])


# ------------------------------------------------------------------------------------------------
# lattice datasets (8-bit images through ToTensor + Normalize(0.5, 0.5), utils/data.py:43-52): two-product mode
# ------------------------------------------------------------------------------------------------
def pixel_images(n, d, g):
    px = torch.randint(0, 256, (n, d), generator=g, dtype=torch.uint8)
    return (px.float() / 255 - 0.5) / 0.5                      # what the reference's transform pipeline yields


def test_lattice_detection(backend):
    from pdm_b200 import EmpiricalDataset
    from pdm_b200.engine import detect_lattice_scale
    g = syn.gen(21)
    dev = backend.device
    img = pixel_images(300, 192, g)
    s = detect_lattice_scale(backend, img.to(dev))
    assert s == 2040.0                                          # 255 * 2^3: y*s = 8*(2p - 255), |.| <= 2040
    assert detect_lattice_scale(backend, (img * 0.5 + 0.5).to(dev)) == 2040.0      # ToTensor only: y*s = 8p
    assert detect_lattice_scale(backend, (torch.randint(0, 2, (50, 64), generator=g).float()).to(dev)) == 2048.0   # dyadic
    assert detect_lattice_scale(backend, (torch.rand(300, 192, generator=g) * 2 - 1).to(dev)) == 0.0
    one_off = img.clone()
    one_off[17, 5] += 1e-4                                      # a single off-lattice value disqualifies the set
    assert detect_lattice_scale(backend, one_off.to(dev)) == 0.0
    ds = EmpiricalDataset(img, backend=backend)
    assert ds.lattice_scale == 2040.0 and ds.scale == 2040.0
    hi, lo = ds.split()
    assert torch.equal(hi.cpu().float()[:, :192], torch.round(img * 2040.0))
    assert lo.cpu().float().abs().max() < 1e-3                  # nothing but fp32 rounding residue


@pytest.mark.parametrize("cta_group", [1, 2])
def test_tensor_stats_lattice_two_products(backend, cta_group):
    """f16x2 (q_hi.y_hi + q_lo.y_hi) on 8-bit image data: same tolerance contract as f16x3."""
    g = syn.gen(22)
    for (m, n, d) in ((200, 700, 192), (300, 1000, 3072)):
        data = pixel_images(n, d, g)
        x0 = data[torch.randint(0, n, (m,), generator=g)]
        t_rows = torch.logspace(-3, 3, m)
        xq = x0 + t_rows.sqrt()[:, None] * torch.randn(m, d, generator=g)
        aux = torch.rand(n, generator=g) * 0.1
        ref = oracle_rows(xq, data, t_rows, aux=aux)
        dev = backend.device
        y = data.to(dev)
        x = xq.to(dev)
        inv_t = (1.0 / t_rows).to(dev)
        ys = backend.prepare_rows(y, n, fixed_scale=2040.0, want_norms=False)
        prep = backend.prepare_rows(x, m)
        for splits, grp in ((0, 0), (3, 2)):
            parts = backend.posterior_stats(precision="f16x2", M=m, N=n, d=d, q_norm=prep["norms"],
                                            y_norm=backend.row_norms(y), inv_temp=inv_t,
                                            q_split=(prep["hi"], prep["lo"], prep["inv_scale"]), y_split=(ys["hi"], None),
                                            y_inv_scale=1.0 / 2040.0, y_aux=aux.to(dev), cta_group=cta_group,
                                            n_splits=splits, m_group=grp)
            out, argmin = backend.merge(parts, inv_t, n)
            check_stats(out.cpu(), argmin.cpu(), ref, aux=True, what=f"f16x2 cg={cta_group} ({m},{n},{d}) S={splits}")


def test_engine_picks_two_products_for_images(backend):
    from pdm_b200 import EmpiricalDataset, PosteriorEngine, EngineConfig, PdmError
    g = syn.gen(23)
    n, d, m = 600, 3 * 16 * 16, 96
    data = pixel_images(n, d, g).view(n, 3, 16, 16)
    eng = PosteriorEngine(EmpiricalDataset(data, backend=backend), EngineConfig())
    assert eng.precision() == "f16x2"
    t_rows = torch.logspace(-2, 2, m)
    xq = data[:m].reshape(m, -1) + t_rows.sqrt()[:, None] * torch.randn(m, d, generator=g)
    ref = oracle_rows(xq, data.reshape(n, -1), t_rows)
    st = eng.stats(xq, t_rows)
    arbitrated_close(st["entropy"], ref["f32"]["entropy"], ref["f64"]["entropy"], atol=2e-5, floor=2 * ref["floor_e"],
                     what="engine f16x2 entropy")
    agree = ref["f32"]["argmin"] == ref["f64"]["argmin"]
    assert torch.equal(st["argmin"].cpu()[agree], ref["f64"]["argmin"][agree])
    # posterior mean: weights_hi.Y + weights_lo.Y only
    got = eng.posterior_mean(xq, t_rows).cpu()
    e64 = 0.5 * orc.pairwise_sqdist(xq.double(), data.reshape(n, -1).double())
    p64 = torch.softmax(-(e64 - e64.min(1, keepdim=True).values) / t_rows.double()[:, None], dim=1)
    ref_mean = p64 @ data.reshape(n, -1).double()
    e32 = 0.5 * orc.pairwise_sqdist(xq, data.reshape(n, -1))
    p32 = torch.softmax(-(e32 - e32.min(1, keepdim=True).values) / t_rows[:, None], dim=1)
    arbitrated_close(got, p32 @ data.reshape(n, -1), ref_mean, atol=2e-5, what="engine f16x2 posterior mean")
    # forcing the two-product mode on continuous data is an error, not a silent loss of accuracy
    cont = EmpiricalDataset(torch.rand(300, 256, generator=g) * 2 - 1, backend=backend)
    with pytest.raises(PdmError):
        PosteriorEngine(cont, EngineConfig(precision="f16x2")).precision()


# ------------------------------------------------------------------------------------------------
# noise regenerated in-kernel from torch's Philox stream
# ------------------------------------------------------------------------------------------------
def test_philox_rows_bit_identical_to_torch_randn(backend):
    """pdm_noised_rows_philox writes exactly torch.randn(...) * sqrt(T) + x0 (utils/stats.py:74, :273), draw after draw."""
    from pdm_b200 import PosteriorEngine
    dev = backend.device
    gen = PosteriorEngine._cuda_generator(dev)
    for shape in ((1024, 3072), (256, 3, 32, 32), (7, 333), (2, 8)):
        torch.manual_seed(77)
        torch.rand(5, device=dev)                                # move the stream off offset 0
        x0 = torch.rand(*shape, device=dev) * 2 - 1
        temps = torch.logspace(-3, 3, 5, device=dev)
        seed, base = gen.initial_seed(), gen.get_offset()
        want = torch.stack([torch.randn(*shape, device=dev) * t.sqrt() + x0 for t in temps])
        after = gen.get_offset()
        step = PosteriorEngine._randn_offset_step(tuple(shape), dev)
        assert after == base + len(temps) * step
        b = shape[0]
        x0f = x0.reshape(b, -1).contiguous()
        got = backend.noised_rows_philox(seed, base, step, x0f, temps.sqrt().contiguous(), want_x=True,
                                         want_split=x0f.shape[1] % 8 == 0)
        assert torch.equal(got["x"].view(len(temps), *shape), want), f"shape {shape}"
        if got["hi"] is not None:
            v = want.reshape(len(temps) * b, -1).double()
            rec = (got["hi"].double() + got["lo"].double()) * got["inv_scale"].double()[:, None]
            amax = v.abs().max(1).values[:, None]
            assert ((rec - v).abs() <= amax * 2.0 ** -20).all()              # bound-scaled split: >= 21 bits of the row max
            assert (got["hi"].float().abs().max() <= 4096)
            nref = (rec ** 2).sum(1)
            assert ((got["norms"].double() - nref).abs() <= 1.2e-7 * nref).all()


def test_fused_noise_engine_matches_unfused(backend):
    from pdm_b200 import EmpiricalDataset, PosteriorEngine, EngineConfig
    g = syn.gen(31)
    n, b, d = 3000, 96, 512
    data = torch.rand(n, d, generator=g) * 2 - 1
    x0 = data[:b].clone()
    temps = torch.logspace(-3, 3, 23)
    ds = EmpiricalDataset(data, backend=backend)
    dev = backend.device
    gen = PosteriorEngine._cuda_generator(dev)
    gen = PosteriorEngine._cuda_generator(dev)
    res = {}
    for fused in (True, False):
        cfg = EngineConfig(max_query_bytes=8 * b * d * 12)          # a few temperatures per block
        cfg.fused_noise = fused
        eng = PosteriorEngine(ds, cfg)
        torch.manual_seed(5)
        res[fused] = (eng.noised_stats(x0, temps), gen.get_offset())
    assert PosteriorEngine._FUSED_NOISE_OK.get(str(dev)) is True, "in-kernel Philox path failed its bit-identity self-check"
    (sf, of), (su, ou) = res[True], res[False]
    assert of == ou                                               # the generator ends where torch.randn would leave it
    assert torch.equal(sf["argmin"], su["argmin"])
    xn = (x0.double() ** 2).sum(1)[None, :] + d * temps.double()[:, None]
    floor = 8 * 2.0 ** -24 * (xn + (data.double() ** 2).sum(1).max()) / temps.double()[:, None]
    for k in ("log_l", "mean_e", "entropy"):
        err = (sf[k].double() - su[k].double()).abs().cpu()
        assert (err <= torch.maximum(1e-4 * su[k].double().abs().cpu() + 2e-5, floor)).all(), k


# ------------------------------------------------------------------------------------------------
# edge cases: ties, tiny / ragged shapes, empty inputs, extreme temperatures
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["exact", "f16x3"])
def test_argmin_ties_take_the_first_index(backend, precision):
    """Duplicated training points: torch.min / argmin return the first index on ties -- across column
    tiles, splits and record merges."""
    g = syn.gen(41)
    base = torch.rand(40, 320, generator=g) * 2 - 1
    data = torch.cat([base, torch.rand(700, 320, generator=g) * 2 - 1, base, base[:7]])     # copies 740.. and 780..
    x = base[:25] + 1e-3 * torch.randn(25, 320, generator=g)
    t_rows = torch.full((25,), 1e-2)
    for splits in (0, 1, 3):
        out, argmin = run_stats(backend, x, data, t_rows, precision, n_splits=splits)
        assert torch.equal(argmin, torch.arange(25)), f"{precision} S={splits}"
        ref = oracle_rows(x, data, t_rows)
        assert torch.equal(ref["f32"]["argmin"], torch.arange(25))
        check_stats(out, argmin, ref, what=f"ties {precision} S={splits}")


def test_tiny_and_ragged_shapes(backend):
    from pdm_b200 import EmpiricalDataset, PosteriorEngine, EngineConfig
    g = syn.gen(42)
    for (n, d, m, prec) in ((1, 7, 3, "exact"), (5, 1, 9, "exact"), (3, 257, 2, "f16x3"), (130, 264, 1, "f16x3"),
                            (257, 1000, 129, "f16x3")):
        data = torch.randn(n, d, generator=g)
        x = data[torch.randint(0, n, (m,), generator=g)] + 0.3 * torch.randn(m, d, generator=g)
        t_rows = torch.logspace(-2, 2, m)
        eng = PosteriorEngine(EmpiricalDataset(data, backend=backend), EngineConfig(precision=prec))
        st = eng.stats(x, t_rows)
        ref = oracle_rows(x, data, t_rows)
        arbitrated_close(st["entropy"], ref["f32"]["entropy"], ref["f64"]["entropy"], atol=2e-5, floor=2 * ref["floor_e"],
                         what=f"entropy ({n},{d},{m}) {prec}")
        agree = ref["f32"]["argmin"] == ref["f64"]["argmin"]
        assert torch.equal(st["argmin"].cpu()[agree], ref["f64"]["argmin"][agree])
        got = eng.posterior_mean(x, t_rows).cpu()
        e64 = 0.5 * orc.pairwise_sqdist(x.double(), data.double())
        p64 = torch.softmax(-(e64 - e64.min(1, keepdim=True).values) / t_rows.double()[:, None], dim=1)
        e32 = 0.5 * orc.pairwise_sqdist(x, data)
        p32 = torch.softmax(-(e32 - e32.min(1, keepdim=True).values) / t_rows[:, None], dim=1)
        arbitrated_close(got, p32 @ data, p64 @ data.double(), atol=2e-5, what=f"mean ({n},{d},{m}) {prec}")


def test_empty_queries_and_extreme_temperatures(backend):
    from pdm_b200 import EmpiricalDataset, PosteriorEngine, EngineConfig
    g = syn.gen(43)
    data = torch.rand(300, 512, generator=g) * 2 - 1
    for prec in ("exact", "f16x3"):
        eng = PosteriorEngine(EmpiricalDataset(data, backend=backend), EngineConfig(precision=prec))
        st = eng.stats(data[:0], torch.ones(0))
        assert st["entropy"].shape == (0,) and st["argmin"].shape == (0,)
        assert eng.posterior_mean(data[:0], torch.ones(0)).shape == (0, 512)
        assert eng.noised_stats(data[:0], torch.tensor([0.1, 1.0]))["entropy"].shape == (2, 0)
        x = data[:6] + 0.05 * torch.randn(6, 512, generator=g)
        t_rows = torch.tensor([1e-30, 1e-12, 1e-6, 1e6, 1e12, 1e30])
        st = eng.stats(x, t_rows)
        for k in ("log_l", "mean_e", "var_e", "entropy", "e_min"):
            assert torch.isfinite(st[k]).all(), f"{prec} {k} not finite at extreme T: {st[k]}"
        assert torch.equal(st["argmin"].cpu(), torch.arange(6))
        # T -> 0: a delta on the nearest point; T -> inf: uniform weights
        assert st["log_l"][:3].abs().max().item() < 1e-6 and st["mean_e"][:3].abs().max().item() < 1e-6
        assert abs(st["entropy"][-1].item()) < 1e-5 and abs(st["log_l"][-1].item() - math.log(300)) < 1e-5
        mean = eng.posterior_mean(x, t_rows).cpu()
        assert torch.isfinite(mean).all()
        torch.testing.assert_close(mean[:3], data[:3], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(mean[-1], data.mean(0), rtol=1e-4, atol=1e-5)


def test_topk_smallest(backend):
    """pdm_topk_smallest_f32 against a stable sort (ties by lower index), incl. k > number of finite entries."""
    g = syn.gen(51)
    dev = backend.device
    for rows, n, k in ((37, 1000, 6), (5, 50_000, 6), (3, 9, 9), (4, 300, 1)):
        x = torch.randn(rows, n, generator=g)
        x[:, n // 3] = x[:, n // 2]                                  # an exact tie in every row
        x[0, :] = x[0, 0]                                            # a constant row: indices 0..k-1
        vals, idx = backend.topk_smallest(x.to(dev), k)
        order = torch.sort(x, dim=1, stable=True)
        assert torch.equal(vals.cpu(), order.values[:, :k]) and torch.equal(idx.cpu(), order.indices[:, :k])
    x = torch.full((2, 5), float("inf"))
    x[0, 3] = 1.0
    vals, idx = backend.topk_smallest(x.to(dev), 3)
    assert idx.cpu()[0].tolist() == [3, 0, 1] and vals.cpu()[0, 0] == 1.0      # +inf entries are still entries (index order)


@pytest.mark.parametrize("kind,n,d,m,k", [("uniform", 5000, 3072, 300, 6), ("pixels", 3000, 1024, 200, 6), ("ragged", 777, 200, 131, 3),
                                          ("tiny", 5, 64, 7, 8), ("offset", 1500, 512, 64, 8)])
def test_topk_epilogue_of_the_fused_pass(backend, kind, n, d, m, k):
    """PosteriorEngine.nearest on the tensor path: the k nearest dataset rows come out of the fused kernel's epilogue (registers
    -> one record per (row, split, column half) -> pdm_topk_merge), no M x N distance tile.  Unrefined it must reproduce the
    dense tile + pdm_topk_smallest_f32 bit for bit (same distances, same tie rule); refined (direct fp64 re-evaluation of the
    8 candidates) it must match a brute-force fp64 search: indices wherever the gap is resolvable, values to 1e-6."""
    from pdm_b200 import EmpiricalDataset, PosteriorEngine, EngineConfig
    g = syn.gen(300 + n)
    if kind == "pixels":
        data = (torch.randint(0, 256, (n, d), generator=g, dtype=torch.uint8).float() / 255 - 0.5) / 0.5
    else:
        data = torch.rand(n, d, generator=g) * 2 - 1
    if n > 100:
        data[n // 2] = data[17]                                      # an exact duplicate: ties resolve to the lower index
    x = data[torch.randint(0, n, (m,), generator=g)] + 0.05 * torch.randn(m, d, generator=g)
    x[0] = data[17 % n]                                              # a query ON the duplicated point
    off = 4000 if kind == "offset" else 0
    ds = EmpiricalDataset(data, backend=backend, index_offset=off, n_total=n + off)
    eng = PosteriorEngine(ds, EngineConfig())
    assert eng.precision() == ("f16x2" if kind == "pixels" else "f16x3")
    launches = backend.launches
    v_raw, i_raw = eng.nearest(x, k, refine=False)
    assert backend.launches - launches <= 4                          # prepare, fused pass, merge: no dense tile, no selection pass
    dense = eng.pairwise_sqdist(x)
    kk = min(k, n)
    v_sel, i_sel = backend.topk_smallest(dense, kk)
    assert torch.equal(v_raw[:, :kk], v_sel) and torch.equal(i_raw[:, :kk], i_sel + off)
    if kk < k:
        assert bool((i_raw[:, kk:] == -1).all()) and bool(torch.isinf(v_raw[:, kk:]).all())
    if n > 100:
        assert i_raw[0, 0].item() == 17 + off and i_raw[0, 1].item() == n // 2 + off
    v, i = eng.nearest(x, k)
    d64 = torch.cdist(x.double(), data.double()).pow(2)
    want_v, want_i = torch.sort(d64, dim=1, stable=True)
    want_v, want_i = want_v[:, :kk], want_i[:, :kk]
    torch.testing.assert_close(v.cpu()[:, :kk].double(), want_v, rtol=2e-6, atol=1e-9)
    floor = 8 * 2.0 ** -24 * ((x.double() ** 2).sum(1, keepdim=True) + (data.double() ** 2).sum(1).max())
    gaps_ok = torch.ones(m, kk, dtype=torch.bool)
    srt = torch.sort(d64, dim=1).values[:, :min(n, 9)]
    for q in range(kk):                                              # slot q is unambiguous when its neighbours in the sorted
        lo_gap = srt[:, q] - srt[:, q - 1] if q > 0 else torch.full((m,), float("inf"), dtype=torch.float64)    # list are far
        hi_gap = srt[:, q + 1] - srt[:, q] if q + 1 < srt.shape[1] else torch.full((m,), float("inf"), dtype=torch.float64)
        gaps_ok[:, q] = (lo_gap > 1e-9) & (hi_gap > 1e-9)
    # selection ran in fp32: candidates beyond slot 8 of the true order could only enter through a near-tie at the cut
    sure = gaps_ok & (srt[:, min(n, 9) - 1:min(n, 9)] - srt[:, :kk] > 2 * floor if n >= 9 else torch.ones(m, kk, dtype=torch.bool))
    assert torch.equal(i.cpu()[:, :kk][sure] - off, want_i[sure])


@pytest.mark.parametrize("precision", ["exact", "f16x3"])
def test_posterior_mean_delta_rows_shortcut(backend, precision):
    """Rows whose posterior is a delta to fp32 resolution are gathered instead of contracted: same result as the full
    weights + contraction path, for all-delta, no-delta and mixed batches."""
    from pdm_b200 import EmpiricalDataset, PosteriorEngine, EngineConfig
    g = syn.gen(61)
    n, d, m = 900, 384, 150
    data = torch.rand(n, d, generator=g) * 2 - 1
    ds = EmpiricalDataset(data, backend=backend)
    x = data[torch.randint(0, n, (m,), generator=g)] + 0.05 * torch.randn(m, d, generator=g)
    for name, temps in (("all delta", torch.full((m,), 1e-3)), ("none", torch.full((m,), 50.0)),
                        ("mixed", torch.logspace(-3, 2.5, m))):
        res = {}
        for on in (True, False):
            cfg = EngineConfig(precision=precision)
            cfg.delta_shortcut = on
            res[on] = PosteriorEngine(ds, cfg).posterior_mean(x, temps).cpu()
        assert (res[True] - res[False]).abs().max().item() <= 3e-7 * data.abs().max().item() + 1e-7, name
        e64 = 0.5 * orc.pairwise_sqdist(x.double(), data.double())
        p64 = torch.softmax(-(e64 - e64.min(1, keepdim=True).values) / temps.double()[:, None], dim=1)
        e32 = 0.5 * orc.pairwise_sqdist(x, data)
        p32 = torch.softmax(-(e32 - e32.min(1, keepdim=True).values) / temps[:, None], dim=1)
        arbitrated_close(res[True], p32 @ data, p64 @ data.double(), atol=2e-5, what=f"delta shortcut {name} {precision}")


def test_random_shapes_against_oracle(backend):
    """Seeded sweep over ragged shapes (every tile boundary: rows vs 128/256, columns vs 128/256, d vs 8/64), continuous
    and 8-bit data, statistics + posterior mean against the fp64 oracle."""
    from pdm_b200 import EmpiricalDataset, PosteriorEngine, EngineConfig
    g = syn.gen(71)
    for case in range(16):
        n = int(torch.randint(1, 900, (1,), generator=g))
        d = int(torch.randint(64, 700, (1,), generator=g))
        m = int(torch.randint(1, 400, (1,), generator=g))
        pixels = case % 3 == 0
        data = pixel_images(n, d, g) if pixels else torch.randn(n, d, generator=g) * 0.7
        x = data[torch.randint(0, n, (m,), generator=g)] + 0.4 * torch.randn(m, d, generator=g)
        t_rows = 10 ** (torch.rand(m, generator=g) * 5 - 2.5)
        eng = PosteriorEngine(EmpiricalDataset(data, backend=backend), EngineConfig())
        assert eng.precision() == ("f16x2" if pixels else "f16x3")
        st = eng.stats(x, t_rows)
        ref = oracle_rows(x, data, t_rows)
        what = f"case {case} (n={n}, d={d}, m={m}, {'pixels' if pixels else 'continuous'})"
        arbitrated_close(st["entropy"], ref["f32"]["entropy"], ref["f64"]["entropy"], atol=2e-5, floor=2 * ref["floor_e"],
                         what=what + " entropy")
        arbitrated_close(st["e_min"], ref["f32"]["e_min"], ref["f64"]["e_min"], atol=1e-5, floor=ref["floor_E"], what=what + " e_min")
        fe2 = ref["floor_e"] * (1 + 2 * ref["f64"]["mean_e"].double())
        arbitrated_close(st["var_e"], ref["f32"]["var_e"], ref["f64"]["var_e"], atol=2e-5, floor=fe2, what=what + " var_e")
        agree = ref["f32"]["argmin"] == ref["f64"]["argmin"]
        assert torch.equal(st["argmin"].cpu()[agree], ref["f64"]["argmin"][agree]), what
        got = eng.posterior_mean(x, t_rows).cpu()
        e64 = 0.5 * orc.pairwise_sqdist(x.double(), data.double())
        p64 = torch.softmax(-(e64 - e64.min(1, keepdim=True).values) / t_rows.double()[:, None], dim=1)
        e32 = 0.5 * orc.pairwise_sqdist(x, data)
        p32 = torch.softmax(-(e32 - e32.min(1, keepdim=True).values) / t_rows[:, None], dim=1)
        arbitrated_close(got, p32 @ data, p64 @ data.double(), atol=2e-5, what=what + " posterior mean")
