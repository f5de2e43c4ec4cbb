"""GPU tests of the certified-delta-posterior path (EngineConfig.screen, include/pdm_b200.h: pdm_screen_*).

The screened engine must give what the unscreened one gives, i.e. stay inside the tolerance contract of
tests/test_gpu_kernels.py against the oracle (fp32 restatement of the reference, fp64 arbiter), on every row --
those it certifies (closed form) and those it sends through the full-precision pass over a row-tile list."""
import math

import pytest
import torch

from oracle import synthetic as syn
from oracle.posterior import pairwise_sqdist as orc_sqdist
from test_gpu_kernels import check_stats, oracle_rows

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def backend(cuda_device):
    from pdm_b200.backend import CudaBackend
    return CudaBackend(cuda_device)


def engines(backend, data, block_rows, d, **cfg_kw):
    from pdm_b200 import EmpiricalDataset, EngineConfig, PosteriorEngine
    ds = EmpiricalDataset(data, backend=backend)
    out = []
    for screen in (True, False):
        cfg = EngineConfig(screen=screen, **cfg_kw)
        cfg.max_query_bytes = block_rows * d * 12          # rows_per_block() == block_rows
        out.append(PosteriorEngine(ds, cfg))
    return out


def stacked(res):
    from pdm_b200.engine import STAT_KEYS
    out = torch.stack([res[k].reshape(-1) for k in STAT_KEYS]).cpu()
    return out, res["argmin"].reshape(-1).cpu()


def run_both(backend, data, x0, temp, block_temps, aux=None, **cfg_kw):
    b, d = x0.shape[0], data[0].numel()
    noise = torch.randn(len(temp), b, d, generator=syn.gen(11))
    scr, ref = engines(backend, data, block_temps * b, d, **cfg_kw)
    fn = lambda i: noise[i].to(backend.device)                               # noqa: E731
    auxd = None if aux is None else aux.to(backend.device)
    r_s = scr.noised_stats(x0, temp, aux=auxd, noise_fn=fn)
    r_u = ref.noised_stats(x0, temp, aux=auxd, noise_fn=fn)
    torch.cuda.synchronize()
    xq = (noise * temp.sqrt()[:, None, None] + x0.reshape(b, -1)[None]).reshape(len(temp) * b, d)
    orc_ref = oracle_rows(xq, data.reshape(len(data), -1), temp.repeat_interleave(b), aux=aux)
    return scr, stacked(r_s), stacked(r_u), orc_ref


def test_screened_schedule_matches_oracle_and_unscreened(backend):
    from pdm_b200 import _cabi as cabi
    n, d, b = 3000, 512, 300                                  # b = 300: row tiles of 256 straddle temperatures
    data = torch.rand(n, d, generator=syn.gen(1)) * 2 - 1
    x0 = data[:b].clone()
    temp = torch.tensor([1e-4, 1e-2, 0.1, 0.3, 1.0, 2.0, 4.0, 8.0, 30.0, 300.0, 3000.0, 1e5])
    scr, (o_s, a_s), (o_u, a_u), ref = run_both(backend, data, x0, temp, block_temps=2)
    check_stats(o_s, a_s, ref, what="screened")
    check_stats(o_u, a_u, ref, what="unscreened")
    rep = scr.screen_report
    assert rep["rows_certified"] >= 4 * b, rep                 # the low-noise temperatures are proven deltas
    assert rep["tiles_full_pass"] < rep["tiles_screened"], rep
    assert rep["rows_unscreened"] >= 4 * b, rep                # the policy stops screening above the first failures
    assert rep["rows_screened"] + rep["rows_unscreened"] == len(temp) * b
    # certified rows carry the closed form; their E_min is the full-precision value, their arg-min the same point
    cert = (o_s[cabi.OUT_L] == 1.0) & (o_s[cabi.OUT_MEAN_E] == 0.0) & (o_s[cabi.OUT_LOG_L] == 0.0)
    assert int(cert.sum()) >= rep["rows_certified"]
    assert torch.equal(a_s, a_u)
    e_s, e_u = o_s[cabi.OUT_E_MIN].double(), o_u[cabi.OUT_E_MIN].double()
    assert ((e_s - e_u).abs() <= ref["floor_E"]).all()
    # and the unscreened engine agrees that those rows are deltas to fp32 resolution
    assert (o_u[cabi.OUT_L][cert] - 1.0).abs().max() <= 2.0 ** -22
    assert o_u[cabi.OUT_MEAN_E][cert].abs().max() <= 1e-5


def test_screening_never_certifies_ties_or_close_pairs(backend):
    from pdm_b200 import _cabi as cabi
    n, d, b = 1500, 256, 64
    g = syn.gen(2)
    data = torch.rand(n, d, generator=g) * 2 - 1
    data[700] = data[3]                                         # exact duplicate: two points share the minimum
    data[900] = data[5] + 1e-3 * torch.randn(d, generator=g)    # near duplicate: gap ~ 1e-4
    data[1100] = data[8] + 0.05 * torch.randn(d, generator=g)   # close pair: gap ~ 0.3
    x0 = data[:b].clone()
    temp = torch.tensor([1e-6, 1e-3, 1e-2, 0.1, 1.0])
    scr, (o_s, a_s), (o_u, a_u), ref = run_both(backend, data, x0, temp, block_temps=5)
    check_stats(o_s, a_s, ref, what="screened, ties")
    assert torch.equal(a_s, a_u)
    l = o_s[cabi.OUT_L].view(len(temp), b)
    assert (l[:, 3] >= 1.5).all()                               # the duplicated point keeps l = 2 at low T: not a delta
    assert a_s.view(len(temp), b)[0, 3] == 3                    # first index on ties, as torch.min
    for k in (cabi.OUT_LOG_L, cabi.OUT_MEAN_E, cabi.OUT_ENTROPY):
        assert torch.allclose(o_s[k], o_u[k], rtol=1e-4, atol=2e-5), k


def test_screened_pixel_lattice_and_aux(backend):
    n, d, b = 2048, 768, 128
    px = torch.randint(0, 256, (n, d), generator=syn.gen(3), dtype=torch.uint8)
    data = (px.float() / 255 - 0.5) / 0.5                       # ToTensor + Normalize(0.5, 0.5): utils/data.py:43-52
    aux = torch.rand(n, generator=syn.gen(4)) + 0.5
    x0 = data[100:100 + b].clone()
    temp = torch.tensor([1e-3, 0.05, 0.5, 2.0, 20.0, 2000.0])
    scr, (o_s, a_s), (o_u, a_u), ref = run_both(backend, data, x0, temp, block_temps=3, aux=aux)
    assert scr.precision() == "f16x2"
    check_stats(o_s, a_s, ref, aux=True, what="screened lattice + aux")
    assert scr.screen_report["rows_certified"] >= 2 * b, scr.screen_report
    assert torch.equal(a_s, a_u)


def test_row_tile_list_launch_is_the_full_launch_restricted(backend):
    """ABI level: with the schedule pinned, the records of the listed row tiles are bit-identical to a full launch."""
    from pdm_b200.engine import pow2_scale_for
    dev = backend.device
    m, n, d = 1000, 1300, 320
    g = syn.gen(5)
    y = (torch.rand(n, d, generator=g) * 2 - 1).to(dev)
    x = (torch.randn(m, d, generator=g)).to(dev)
    inv_t = (1.0 / (torch.rand(m, generator=g) * 50 + 0.5)).to(dev)
    scale = pow2_scale_for(float(backend.absmax(y).item()))
    ys = backend.prepare_rows(y, n, fixed_scale=scale, want_norms=False)
    prep = backend.prepare_rows(x, m)
    kw = dict(precision="f16x3", M=m, N=n, d=d, q_norm=prep["norms"], y_norm=backend.row_norms(y), inv_temp=inv_t,
              q_split=(prep["hi"], prep["lo"], prep["inv_scale"]), y_split=(ys["hi"], ys["lo"]), y_inv_scale=1.0 / scale,
              n_splits=3, m_group=2, cta_group=2)
    full = backend.posterior_stats(**kw)
    tiles = torch.tensor([3, 1], dtype=torch.int32, device=dev)             # any order; tile 3 is ragged (rows 768..999)
    part = backend.posterior_stats(row_tiles=tiles, n_row_tiles=2, **kw)
    torch.cuda.synchronize()
    assert full.shape == part.shape
    for t in (1, 3):
        r0, r1 = t * 256, min(m, (t + 1) * 256)
        assert torch.equal(full[:, r0:r1].contiguous().view(torch.int32), part[:, r0:r1].contiguous().view(torch.int32)), t


def test_screened_cifar_shape_block(backend):
    """C2 shape (N = 50 000, d = 3072), a slice of the DDPM schedule around the certification boundary."""
    from pdm_b200 import _cabi as cabi
    n, d, b = 50_000, 3072, 256
    data = torch.rand(n, d, generator=syn.gen(6)) * 2 - 1
    x0 = data[:b].clone()
    temp = torch.tensor([1e-4, 0.5, 4.0, 8.0, 12.0, 40.0, 1000.0])
    noise = torch.randn(len(temp), b, d, generator=syn.gen(12))
    scr, ref = engines(backend, data, 3 * b, d)
    fn = lambda i: noise[i].to(backend.device)                               # noqa: E731
    o_s, a_s = stacked(scr.noised_stats(x0, temp, noise_fn=fn))
    o_u, a_u = stacked(ref.noised_stats(x0, temp, noise_fn=fn))
    assert torch.equal(a_s, a_u)
    rep = scr.screen_report
    assert rep["rows_certified"] >= 3 * b, rep                  # T <= 4 is certified at this shape
    t_rows = temp.repeat_interleave(b).double()
    xn = (noise.double() * temp.double().sqrt()[:, None, None] + x0.double()[None]).pow(2).sum(-1).reshape(-1)
    floor_e = 8 * 2.0 ** -24 * (xn + float(d)) / t_rows
    for k, atol in ((cabi.OUT_LOG_L, 1e-5), (cabi.OUT_MEAN_E, 1e-5), (cabi.OUT_ENTROPY, 2e-5)):
        err = (o_s[k].double() - o_u[k].double()).abs()
        tol = 1e-4 * o_u[k].double().abs() + atol + 2 * floor_e
        assert (err <= tol).all(), (k, err.max().item())
    err_m = (o_s[cabi.OUT_E_MIN].double() - o_u[cabi.OUT_E_MIN].double()).abs()
    assert (err_m <= 8 * 2.0 ** -24 * (xn + float(d))).all()


def test_screened_descending_schedule(backend):
    """A schedule that starts at the high-noise end: failed attempts back off, the low-noise end is still certified."""
    n, d, b = 2000, 256, 256
    data = torch.rand(n, d, generator=syn.gen(7)) * 2 - 1
    x0 = data[500:500 + b].clone()
    temp = torch.tensor([1e4, 1e3, 300.0, 100.0, 30.0, 10.0, 0.3, 1e-2, 1e-4])
    scr, (o_s, a_s), (o_u, a_u), ref = run_both(backend, data, x0, temp, block_temps=1)
    check_stats(o_s, a_s, ref, what="screened, descending")
    rep = scr.screen_report
    assert rep["rows_certified"] >= 2 * b and rep["rows_unscreened"] >= 2 * b, rep
    assert torch.equal(a_s, a_u)


def test_screened_posterior_mean(backend):
    """Ideal denoiser (Scheduler.true_posterior_mean_x0, diffusion/scheduler/scheduler.py:58-69) with screening: low-noise
    blocks are answered by a gather of certified nearest points, high-noise blocks by the two contractions."""
    from pdm_b200 import EmpiricalDataset, EngineConfig, PosteriorEngine
    from oracle import posterior as orc
    n, d, b = 3000, 512, 200
    g = syn.gen(8)
    data = torch.rand(n, d, generator=g) * 2 - 1
    ds = EmpiricalDataset(data, backend=backend)
    scr = PosteriorEngine(ds, EngineConfig(screen=True))
    ref = PosteriorEngine(ds, EngineConfig(screen=False))
    for alpha_bar, expect_certified in ((0.999, True), (0.9, True), (0.02, False), (0.95, True)):
        ab = torch.tensor(alpha_bar)
        xt = ab.sqrt() * data[:b] + (1 - ab).sqrt() * torch.randn(b, d, generator=g)
        t_rows, post = ((1 - ab) / ab).expand(b), ab.rsqrt().expand(b)
        before = scr.screen_report.get("pm_rows_certified", 0)
        got = scr.posterior_mean(xt, t_rows, post=post).cpu()
        scr._pm_poll(wait=True)              # the counts follow the call (no host read inside it): collect them
        plain = ref.posterior_mean(xt, t_rows, post=post).cpu()
        want = orc.posterior_mean_x0(xt, ab, data, dtype=torch.float64)
        assert (scr.screen_report.get("pm_rows_certified", 0) > before) == expect_certified, (alpha_bar, scr.screen_report)
        assert (got.double() - want).abs().max().item() < 1e-4, alpha_bar
        assert (got - plain).abs().max().item() < 1e-4, alpha_bar
    # after the failure at alpha_bar = 0.02 (T = 49) the mark sits at T / 2; T = 0.053 (alpha_bar = 0.95) is screened again
    assert scr.screen_report["pm_rows_screened"] == 4 * b


@pytest.mark.parametrize("seed", [21, 22, 23, 24, 25, 26])
def test_screened_sweep_over_datasets_and_ragged_shapes(backend, seed):
    """Seeded sweep: clustered / heavy-tailed / pixel / duplicated datasets, shapes ragged against every tile size,
    unsorted temperatures.  Every row -- certified or not -- must stay inside the oracle tolerance."""
    g = syn.gen(seed)
    n = int(torch.randint(300, 2600, (1,), generator=g))
    d = int(torch.randint(8, 90, (1,), generator=g)) * 8
    b = int(torch.randint(5, 140, (1,), generator=g))
    kind = seed % 4
    if kind == 0:                                               # tight Gaussian clusters: many close pairs
        centres = torch.randn(7, d, generator=g) * 2
        data = centres[torch.randint(0, 7, (n,), generator=g)] + 0.05 * torch.randn(n, d, generator=g)
    elif kind == 1:                                             # heavy tails: row norms spread over decades
        data = torch.randn(n, d, generator=g) * torch.exp(1.5 * torch.randn(n, 1, generator=g))
    elif kind == 2:                                             # 8-bit pixels with repeated images
        px = torch.randint(0, 256, (n, d), generator=g, dtype=torch.uint8)
        px[n // 2:n // 2 + 20] = px[:20]
        data = (px.float() / 255 - 0.5) / 0.5
    else:                                                       # uniform cube, offset from the origin
        data = torch.rand(n, d, generator=g) * 2 + 3
    x0 = data[torch.randint(0, n, (b,), generator=g)].clone()
    temp = torch.exp(torch.empty(9).uniform_(math.log(1e-5), math.log(1e4), generator=g))
    temp = temp.sort().values if seed % 2 else temp             # odd seeds: ascending; even seeds: as drawn
    scr, (o_s, a_s), (o_u, a_u), ref = run_both(backend, data, x0, temp, block_temps=2)
    check_stats(o_s, a_s, ref, what=f"screened sweep seed {seed} kind {kind} n={n} d={d} b={b}")
    agree = ref["f32"]["argmin"] == ref["f64"]["argmin"]
    assert torch.equal(a_s[agree], a_u[agree])


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_one_product_error_bounds_hold_for_every_pair(backend, kind):
    """The certificates rest on |E_pass - E| <= delta for EVERY pair.  Energy tiles of the fp16 one-product pass and of the
    E4M3 pass against fp64 energies, with kappa = 1 (the engine adds 25 % / 2 % on top), over dataset families."""
    from pdm_b200.engine import pow2_scale_for
    dev = backend.device
    g = syn.gen(40 + kind)
    n, d, m = 1500, 384, 300
    if kind == 0:
        data = torch.rand(n, d, generator=g) * 2 - 1
    elif kind == 1:
        data = torch.randn(n, d, generator=g) * torch.exp(1.5 * torch.randn(n, 1, generator=g))
    elif kind == 2:
        data = (torch.randint(0, 256, (n, d), generator=g, dtype=torch.uint8).float() / 255 - 0.5) / 0.5
    else:
        data = torch.rand(n, d, generator=g) * 2 + 3
    x = data[torch.randint(0, n, (m,), generator=g)] + torch.randn(m, d, generator=g) * torch.logspace(-3, 2, m)[:, None]
    y = data.to(dev).contiguous()
    scale = pow2_scale_for(float(backend.absmax(y).item()))
    ys = backend.prepare_rows(y, n, fixed_scale=scale, want_norms=False)
    prep = backend.prepare_rows(x.to(dev).contiguous(), m)
    y_norm = backend.row_norms(y)
    e64 = 0.5 * orc_sqdist(x.double(), data.double())
    xn, yn_max = (x.double() ** 2).sum(1), (data.double() ** 2).sum(1).max()
    slack = 8 * 2.0 ** -24 * (xn + yn_max)[:, None]                         # fp32 round-off of the norm expansion itself
    common = dict(M=m, N=n, d=d, q_norm=prep["norms"], y_norm=y_norm, inv_temp=None, want_partials=False)
    # fp16 one-product pass: delta = 2^-10 ||x|| max||y||
    e1 = torch.empty(m, n, device=dev)
    backend.posterior_stats(precision="f16x1", q_split=(prep["hi"], prep["lo"], prep["inv_scale"]), y_split=(ys["hi"], ys["lo"]),
                            y_inv_scale=1.0 / scale, energy_out=e1, **common)
    d1 = 2.0 ** -10 * xn.sqrt() * yn_max.sqrt()
    assert ((e1.cpu().double() - e64).abs() <= d1[:, None] + slack).all()
    # E4M3 pass: delta = ex ||y||max + (||x|| + ex) ey from the exact rounding deviations
    q8, q_err = backend.split_to_e4m3(prep["hi"], prep["lo"], d)
    y8, y_err = backend.split_to_e4m3(ys["hi"], ys["lo"], d)
    ex = (q_err * prep["inv_scale"]).cpu().double()
    ey = float(y_err.max().item()) / scale
    e8 = torch.empty(m, n, device=dev)
    backend.posterior_stats(precision="f8x1", q_split=(q8, None, prep["inv_scale"] * 16.0), y_split=(y8, None),
                            y_inv_scale=16.0 / scale, energy_out=e8, **common)
    d8 = ex * yn_max.sqrt() + (xn.sqrt() + ex) * ey
    err8 = (e8.cpu().double() - e64).abs()
    assert (err8 <= d8[:, None] + slack).all(), (err8 / (d8[:, None] + slack)).max()
    assert (err8.max(1).values > 1e-3 * d8).any()          # the pass really ran on 4-bit significands
    # the deviations are what they claim to be
    rec = (prep["hi"].double() + prep["lo"].double())[:, :d].cpu()
    from_bytes = q8.view(torch.float8_e4m3fn).double()[:, :d].cpu() * 16
    assert torch.allclose((rec - from_bytes).norm(dim=1), q_err.cpu().double(), rtol=1e-4, atol=1e-12)


@pytest.mark.parametrize("compact", [True, False])
def test_remembered_boundary_calls_match_the_oracle(backend, compact):
    """Calls two to four on one engine: the certifiable range is remembered, every block runs its cascade and its
    full-precision pass without a host read-back, the E4M3 stage gets its own mark, and (``compact``) the unproven rows of the
    screened range are gathered into dense row tiles.  Each call against the oracle on every row, like the first."""
    from pdm_b200 import EmpiricalDataset, PosteriorEngine, EngineConfig
    n, d, b = 3000, 512, 300                                  # b = 300: row tiles of 256 straddle temperatures
    g = syn.gen(41)
    centres = torch.rand(40, d, generator=g) * 2 - 1          # clustered: rows near the boundary are a mix of proven / unproven
    data = (centres[torch.randint(0, 40, (n,), generator=g)] + 0.25 * torch.randn(n, d, generator=g)).clamp(-1, 1)
    data[7] = data[3]                                         # a duplicate: query 3 never certifies, at any temperature
    x0 = data[:b].clone()
    temp = torch.logspace(-4, 4, 24)
    noise = torch.randn(len(temp), b, d, generator=syn.gen(42))
    fn = lambda i: noise[i].to(backend.device)                # noqa: E731
    cfg = EngineConfig(screen=True, screen_compact=compact)
    cfg.max_query_bytes = 8 * b * d * 12                      # blocks of eight temperatures
    eng = PosteriorEngine(EmpiricalDataset(data, backend=backend), cfg)
    xq = (noise * temp.sqrt()[:, None, None] + x0[None]).reshape(len(temp) * b, d)
    ref = oracle_rows(xq, data, temp.repeat_interleave(b))
    outs = []
    for call in range(4):
        rep0 = dict(eng.screen_report)
        o, a = stacked(eng.noised_stats(x0, temp, noise_fn=fn))
        check_stats(o, a, ref, what=f"call {call} compact={compact}")
        outs.append((o, a, {k: eng.screen_report.get(k, 0) - rep0.get(k, 0) for k in eng.screen_report}))
    assert eng._screen_prior is not None and eng._screen_prior_f8 is not None
    open_rows = outs[3][2]["rows_screened"] - outs[3][2]["rows_certified"]
    assert open_rows >= len(temp) // 4                        # the duplicate's query at every screened temperature, at least
    if compact:
        assert outs[3][2]["tiles_full_pass"] <= -(-open_rows // 256) + 3, outs[3][2]     # dense tiles, rounded up per block
    assert outs[3][2].get("f8_tiles_screened", 0) <= outs[1][2].get("f8_tiles_screened", 0)
    assert torch.equal(outs[3][1], outs[0][1])                # arg-min: the same points on every path
