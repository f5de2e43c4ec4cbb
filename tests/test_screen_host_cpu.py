"""CPU tests of the certified-delta-posterior host logic (pdm_b200.engine: EngineConfig.screen) on a torch test double
that restates the pdm_screen_* contract of include/pdm_b200.h: the screening policy over a schedule, the row-tile list,
the closed form, and -- independently of any kernel -- the SOUNDNESS of the certificate: every row it certifies is
checked in fp64 to have all other weights below exp(-g)."""
import math

import torch

from fake_backend import FakeBackend, PART, _pack_idx
from oracle import posterior as orc
from oracle import synthetic as syn
from pdm_b200 import EmpiricalDataset, EngineConfig, PosteriorEngine


def _split16(v: torch.Tensor, scale: torch.Tensor):
    s = v * scale
    hi = s.half()
    lo = (s - hi.float()).half()
    return hi, lo


class SplitFakeBackend(FakeBackend):
    """FakeBackend + the fp16 hi/lo operand split, the one-/three-product contractions on it, row-tile lists and the
    three screening entry points (torch restatement of csrc/screen.cu)."""
    row_tile = 8

    def supports_tensor_path(self) -> bool:
        return True

    def prepare_rows(self, src, rows, *, noise=None, sigma=None, post=None, fixed_scale=0.0, want_x=False,
                     want_norms=True, want_split=True) -> dict:
        r = super().prepare_rows(src, rows, noise=noise, sigma=sigma, post=post)
        v = r["x"]
        if fixed_scale > 0:
            scale = torch.full((v.shape[0], 1), float(fixed_scale))
        else:
            amax = v.abs().max(1, keepdim=True).values.clamp(min=1e-30)
            scale = 2.0 ** (11 - torch.floor(torch.log2(amax)))            # max|v|*scale in [2^11, 2^12)
        r["hi"], r["lo"] = _split16(v, scale)
        r["inv_scale"] = (1.0 / scale).reshape(-1)
        return r

    def posterior_stats(self, *, precision, M, N, d, q_norm, y_norm, inv_temp, q_split=None, y_split=None,
                        y_inv_scale=1.0, y_aux=None, index_offset=0, n_splits=0, row_tiles=None, n_row_tiles=0,
                        n_row_tiles_dev=None, **kw):
        if precision == "exact":
            return super().posterior_stats(precision=precision, M=M, N=N, d=d, q_norm=q_norm, y_norm=y_norm,
                                           inv_temp=inv_temp, y_aux=y_aux, index_offset=index_offset, n_splits=n_splits, **kw)
        q_hi, q_lo, q_inv = q_split
        y_hi, y_lo = y_split
        if precision == "f8x1":                             # E4M3 bytes; the inverse scales already carry the factor 16
            q_hi, y_hi = q_hi.view(torch.float8_e4m3fn), y_hi.view(torch.float8_e4m3fn)
        qv = q_hi.double() if precision in ("f16x1", "f8x1") else q_hi.double() + q_lo.double()
        yv = y_hi.double() if (precision in ("f16x1", "f16x2", "f8x1") or y_lo is None) else y_hi.double() + y_lo.double()
        q = (qv * q_inv.double()[:, None]).float()
        y = (yv * y_inv_scale).float()
        self.calls.append(f"stats:{precision}:{'list' if row_tiles is not None else 'all'}")
        parts = super().posterior_stats(precision="exact", M=M, N=N, d=d, q_norm=q_norm, y_norm=y_norm, inv_temp=inv_temp,
                                        q=q, y=y, y_aux=y_aux, index_offset=index_offset, n_splits=n_splits,
                                        energy_out=kw.get("energy_out"), energy_mult=kw.get("energy_mult", 1.0))
        if row_tiles is not None:
            keep = torch.zeros(M, dtype=torch.bool)
            if n_row_tiles_dev is not None:                 # ABI v4: the list's length lives "on the device"
                n_row_tiles = min(int(n_row_tiles), int(n_row_tiles_dev[0]))
                self.calls.append(f"tiles:{precision}:{n_row_tiles}")
            for t in row_tiles[:n_row_tiles].tolist():
                keep[t * self.row_tile:(t + 1) * self.row_tile] = True
            parts[~keep] = float("nan")                    # rows outside the list are never written by the kernel
        return parts

    # ---- E4M3 first stage (class attribute ``with_f8`` switches it on) ----
    def split_to_e4m3(self, hi, lo, d):
        v = hi.float() + (lo.float() if lo is not None else 0.0)
        b = (v / 16).to(torch.float8_e4m3fn)
        err = (v.double() - b.double() * 16).norm(dim=1).float() * (1 + 1e-5)
        return b.view(torch.uint8), err

    def screen_temperatures_f8(self, q_norm, q_err, q_inv_scale, inv_temp, y_norm_max, y_err_max, g, e_star, kappa):
        ex = q_err * q_inv_scale
        delta = kappa * (ex * y_norm_max.sqrt() + (q_norm.sqrt() + ex) * y_err_max) + 2.4e-7 * (q_norm + y_norm_max)
        return e_star / (g / inv_temp + 2 * delta)

    def screen_tile_list(self, flags, rows_per_tile):
        tiles = (flags.numel() + rows_per_tile - 1) // rows_per_tile
        listed = [t for t in range(tiles) if not bool(flags[t * rows_per_tile:(t + 1) * rows_per_tile].all())]
        tile_list = torch.zeros(max(1, tiles), dtype=torch.int32)
        tile_list[:len(listed)] = torch.tensor(listed, dtype=torch.int32)
        return tile_list, torch.tensor([len(listed)], dtype=torch.int32)

    # ---- ABI v4: tile lists whose length stays on the device (sync-free posterior mean) ----
    def _listed_rows(self, m, tiles):
        tl, rpt, n_max, n_dev = tiles
        n = min(int(n_max), int(n_dev[0])) if n_dev is not None else int(n_max)
        keep = torch.zeros(m, dtype=torch.bool)
        for t in tl[:n].tolist():
            keep[t * rpt:(t + 1) * rpt] = True
        return keep, n

    def delta_tile_list(self, l, rows_per_tile):
        flags = ((l - 1.0) <= 2.0 ** -23).to(torch.uint8)
        tl, n = self.screen_tile_list(flags, rows_per_tile)
        return flags, tl, n

    def screen_merge_stage(self, tile_list, n_listed, max_tiles, rows_per_tile, flags_b, arg_b, flags, arg):
        keep, _ = self._listed_rows(flags.numel(), (tile_list, rows_per_tile, max_tiles, n_listed))
        take = keep & (flags == 0)
        flags[take] = flags_b[take]
        arg[take] = arg_b[take]

    def gather_rows(self, src, idx, index_offset, flags, out):
        sel = torch.ones(idx.numel(), dtype=torch.bool) if flags is None else flags != 0
        j = idx - index_offset
        own = (j >= 0) & (j < src.shape[0])
        picked = src[j.clamp(0, src.shape[0] - 1)] * own[:, None].to(src.dtype)
        out[sel] = picked[sel]
        return out

    def transpose_split(self, y, scale):
        hi, lo = _split16(y.t().contiguous(), torch.tensor(float(scale)))
        return hi, lo

    def weights_from_energy(self, energy, e_min, l, inv_temp, *, split, tiles=None):
        p = super().weights_from_energy(energy, e_min, l, inv_temp, split=False)
        if tiles is not None:
            keep, n = self._listed_rows(energy.shape[0], tiles)
            self.calls.append(f"weights:{n}")
            p[~keep] = float("nan")                         # rows outside the list are never written
        if not split:
            return p
        return _split16(p, torch.tensor(16384.0))

    def split_gemm(self, a_hi, a_lo, b_hi, b_lo, k, scale, out=None, accumulate=False, cta_group=0, tiles=None):
        a = a_hi.double() + a_lo.double()
        b = b_hi.double() + (b_lo.double() if b_lo is not None else 0.0)
        r = (torch.nan_to_num(a) @ b.t() * scale).float()
        if out is None:
            out = torch.zeros_like(r)
        keep = torch.ones(a.shape[0], dtype=torch.bool)
        if tiles is not None:
            keep, n = self._listed_rows(a.shape[0], (tiles[0], self.row_tile, tiles[1], tiles[2]))
            self.calls.append(f"gemm2:{n}")
        out[keep] = r[keep] if not accumulate else out[keep] + r[keep]
        return out

    def screen_temperatures(self, q_norm, inv_temp, y_norm_max, g, e_star, kappa):
        delta = kappa * 2.0 ** -10 * q_norm.sqrt() * y_norm_max.sqrt() + 2.4e-7 * (q_norm + y_norm_max)
        return e_star / (g / inv_temp + 2 * delta)

    def screen_certify(self, screen_out, e_star, rows_per_tile):
        l = screen_out[7]
        a1 = screen_out[2] * l
        flags = ((l > 0.99) & (l < 1.25) & (a1 >= 0) & (a1 < 0.9 * e_star * math.exp(-e_star))).to(torch.uint8)
        m = flags.numel()
        tiles = (m + rows_per_tile - 1) // rows_per_tile
        listed = [t for t in range(tiles) if not bool(flags[t * rows_per_tile:(t + 1) * rows_per_tile].all())]
        tile_list = torch.zeros(max(1, tiles), dtype=torch.int32)
        tile_list[:len(listed)] = torch.tensor(listed, dtype=torch.int32)
        return flags, tile_list, torch.tensor([len(listed)], dtype=torch.int32)

    def screen_finalize(self, flags, screen_argmin, d, q_split, q_norm, y_split, y_inv_scale, y_norm, y_aux, index_offset,
                        n_local, n_total, out, argmin):
        q_hi, q_lo, q_inv = q_split
        y_hi, y_lo = y_split
        for r in torch.nonzero(flags).flatten().tolist():
            k = int(screen_argmin[r])
            j = k - index_offset
            out[1:5, r] = 0.0
            out[6, r] = -math.log(n_total)
            out[7, r] = 1.0
            argmin[r] = k
            if j < 0 or j >= n_local:
                out[0, r], out[5, r] = math.inf, -math.inf
                continue
            yv = y_hi[j].double() + (y_lo[j].double() if y_lo is not None else 0.0)
            s = ((q_hi[r].double() + q_lo[r].double()) * yv).sum().float()
            u = (s * (-2.0 * q_inv[r] * y_inv_scale) + q_norm[r]) + y_norm[j]
            out[0, r] = 0.5 * u
            out[5, r] = y_aux[j] if y_aux is not None else 0.0


def _setup(n=400, d=96, b=12, seed=1):
    data = torch.rand(n, d, generator=syn.gen(seed)) * 2 - 1
    data[37] = data[5]                                      # a duplicate: the row of query 5 must never be certified
    return data, data[:b].clone()


def _run(data, x0, temp, screen, block_temps, aux=None, f8=False):
    be = SplitFakeBackend()
    cfg = EngineConfig(precision="f16x3", screen=screen, screen_f8=f8)
    cfg.max_query_bytes = block_temps * x0.shape[0] * data.shape[1] * 12
    eng = PosteriorEngine(EmpiricalDataset(data, backend=be), cfg)
    noise = torch.randn(len(temp), x0.shape[0], data.shape[1], generator=syn.gen(9))
    res = eng.noised_stats(x0, temp, aux=aux, noise_fn=lambda i: noise[i])
    return eng, be, res, noise


def test_screened_schedule_equals_unscreened_and_policy_stops():
    data, x0 = _setup()
    temp = torch.tensor([1e-4, 1e-3, 1e-2, 0.05, 0.2, 0.5, 1.0, 3.0, 10.0, 100.0, 1e3, 1e4])
    aux = torch.rand(len(data), generator=syn.gen(3))
    eng_s, be_s, r_s, noise = _run(data, x0, temp, True, 2, aux=aux)
    eng_u, _, r_u, _ = _run(data, x0, temp, False, 2, aux=aux)
    rep = eng_s.screen_report
    assert rep["rows_certified"] > 0 and rep["rows_unscreened"] > 0 and rep["tiles_full_pass"] < rep["tiles_screened"], rep
    assert torch.equal(r_s["argmin"], r_u["argmin"])
    for k in ("log_l", "mean_e", "mean_e2", "var_e", "entropy", "aux_mean", "l"):
        assert torch.isfinite(r_s[k]).all(), k               # nothing of an unlisted tile's NaN records leaks
        assert torch.allclose(r_s[k], r_u[k], rtol=1e-4, atol=2e-6), (k, (r_s[k] - r_u[k]).abs().max())
    xn = ((noise * temp.sqrt()[:, None, None] + x0[None]) ** 2).sum(-1)
    assert ((r_s["e_min"] - r_u["e_min"]).abs() <= 8 * 2.0 ** -24 * (xn + data.shape[1])).all()
    # the one-product pass ran only on the screened blocks, the full pass only on row-tile lists there
    assert be_s.calls.count("stats:f16x1:all") == rep["rows_screened"] // (2 * len(x0))
    assert "stats:f16x3:all" in be_s.calls and "stats:f16x3:list" in be_s.calls
    # the duplicated training point keeps l = 2: its query is never a certified delta
    assert (r_s["l"][:4, 5] >= 1.9).all() and (r_s["argmin"][:4, 5] == 5).all()


def test_repeated_calls_remember_the_certifiable_range():
    """Second and later calls on the same dataset: the boundary found by the first call's probing is remembered, every block
    screens its rows below 1.5x that mark in ONE cascade with the tile list and its length left on the device (no probing, no
    read-back inside the call) -- same results, and a query sitting on a duplicated training point (never certifiable, at any
    temperature) does not drag the boundary down."""
    data, x0 = _setup()
    temp = torch.tensor([1e-4, 1e-3, 1e-2, 0.05, 0.2, 0.5, 1.0, 3.0, 10.0, 100.0, 1e3, 1e4])
    be = SplitFakeBackend()
    cfg = EngineConfig(precision="f16x3", screen=True, screen_f8=False)
    cfg.max_query_bytes = 4 * x0.shape[0] * data.shape[1] * 12            # blocks of four temperatures
    eng = PosteriorEngine(EmpiricalDataset(data, backend=be), cfg)
    noise = torch.randn(len(temp), x0.shape[0], data.shape[1], generator=syn.gen(9))
    first = eng.noised_stats(x0, temp, noise_fn=lambda i: noise[i])
    prior = eng._screen_prior
    assert prior is not None and prior > 0.04, prior                      # NOT 1e-4, where the duplicate's query already fails
    n_probe = len(be.calls)
    second = eng.noised_stats(x0, temp, noise_fn=lambda i: noise[i])
    calls = be.calls[n_probe:]
    assert any(c.startswith("tiles:f16x3:") for c in calls)               # device-side list lengths: the sync-free path
    assert calls.count("stats:f16x1:all") <= 3                            # one cascade per block that reaches below the mark
    for k in ("log_l", "mean_e", "var_e", "entropy", "l"):
        assert torch.isfinite(second[k]).all(), k
        assert torch.allclose(second[k], first[k], rtol=1e-5, atol=1e-6), k
    assert torch.equal(second["argmin"], first["argmin"])
    assert (second["l"][:4, 5] >= 1.9).all()                              # the duplicate's query still goes through the full pass
    plain = PosteriorEngine(EmpiricalDataset(data, backend=SplitFakeBackend()), EngineConfig(precision="f16x3", screen=False))
    ref = plain.noised_stats(x0, temp, noise_fn=lambda i: noise[i])
    for k in ("log_l", "mean_e", "var_e", "entropy"):
        assert torch.allclose(second[k], ref[k], rtol=1e-4, atol=2e-6), k
    # a descending schedule takes the same path (the certifiable rows sit at the END of a block)
    third = eng.noised_stats(x0, temp.flip(0), noise_fn=lambda i: noise[len(temp) - 1 - i])
    assert torch.allclose(third["entropy"].flip(0), ref["entropy"], rtol=1e-4, atol=2e-6)


def test_certificate_is_sound_in_fp64():
    """Every certified row: all other points' weights are below exp(-g) in exact arithmetic."""
    data, x0 = _setup(n=600, d=128, b=16, seed=4)
    data[100] = data[3] + 0.02 * torch.randn(128, generator=syn.gen(5))     # close pairs around the decision boundary
    data[101] = data[4] + 0.2 * torch.randn(128, generator=syn.gen(6))
    temp = torch.logspace(-3, 1.5, 19)
    eng, _, res, noise = _run(data, x0, temp, True, 19)
    g = 17.0 + math.log(len(data))
    xt = (noise.double() * temp.double().sqrt()[:, None, None] + x0.double()[None])
    e = 0.5 * orc.pairwise_sqdist(xt.reshape(-1, data.shape[1]), data.double())
    srt = e.sort(dim=1).values
    gap = ((srt[:, 1] - srt[:, 0]) / temp.double().repeat_interleave(len(x0))).view(len(temp), len(x0))
    cert = (res["l"] == 1.0) & (res["mean_e"] == 0.0) & (res["log_l"] == 0.0) & (res["var_e"] == 0.0)
    assert int(cert.sum()) >= eng.screen_report["rows_certified"] > 0
    strict = cert & (gap < 80)                      # rows whose unscreened result is not itself an exact delta in fp32
    assert (gap[strict] > g).all(), gap[strict].min()
    assert (gap[cert] > g).all()
    # and the test is not vacuous: some rows with a gap above g were left uncertified (the bound is conservative) ...
    assert int(((gap > g) & ~cert).sum()) > 0
    # ... while every row with a gap below g went through the full pass
    assert not bool((cert & (gap <= g)).any())


def test_descending_schedule_backs_off():
    data, x0 = _setup()
    temp = torch.tensor([1e4, 3e3, 1e3, 300.0, 100.0, 30.0, 10.0, 1e-2, 1e-3, 1e-4])
    eng_s, be_s, r_s, _ = _run(data, x0, temp, True, 1)
    eng_u, _, r_u, _ = _run(data, x0, temp, False, 1)
    rep = eng_s.screen_report
    # a failed attempt is repeated only once the temperature has dropped fourfold: 1e4, 1e3, 100, 10 fail; 1e-2.. certify
    assert be_s.calls.count("stats:f16x1:all") <= len(temp) - 3, be_s.calls
    assert rep["rows_certified"] > 0 and rep["rows_unscreened"] >= 3 * len(x0), rep
    for k in ("log_l", "mean_e", "entropy"):
        assert torch.allclose(r_s[k], r_u[k], rtol=1e-4, atol=2e-6), k
    assert torch.equal(r_s["argmin"], r_u["argmin"])


def _dataset(kind: int, n: int, d: int, g) -> torch.Tensor:
    if kind == 0:                                               # tight clusters: many near pairs
        centres = torch.randn(5, d, generator=g) * 2
        return centres[torch.randint(0, 5, (n,), generator=g)] + 0.05 * torch.randn(n, d, generator=g)
    if kind == 1:                                               # heavy tails: row norms over decades
        return torch.randn(n, d, generator=g) * torch.exp(1.5 * torch.randn(n, 1, generator=g))
    if kind == 2:                                               # 8-bit pixels, some repeated
        px = torch.randint(0, 256, (n, d), generator=g, dtype=torch.uint8)
        px[n // 2:n // 2 + 10] = px[:10]
        return (px.float() / 255 - 0.5) / 0.5
    return torch.rand(n, d, generator=g) * 2 + 3                # cube away from the origin


def test_certificate_soundness_sweep():
    """fp64 check of every certified row over dataset families and seeds (no certified row with a gap below g*T)."""
    total_cert = 0
    for seed in range(8):
        g = syn.gen(100 + seed)
        n, d, b = 300 + 37 * seed, 64 + 8 * seed, 10
        data = _dataset(seed % 4, n, d, g)
        x0 = data[torch.randint(0, n, (b,), generator=g)].clone()
        temp = torch.logspace(-5, 3, 17)
        eng, _, res, noise = _run(data, x0, temp, True, 17)
        gthr = 17.0 + math.log(n)
        xt = (noise.double() * temp.double().sqrt()[:, None, None] + x0.double()[None])
        e = 0.5 * orc.pairwise_sqdist(xt.reshape(-1, d), data.double())
        srt = e.sort(dim=1).values
        gap = ((srt[:, 1] - srt[:, 0]) / temp.double().repeat_interleave(b)).view(len(temp), b)
        # rows the engine itself reports as certified: closed form AND counted by the report
        cert = (res["l"] == 1.0) & (res["mean_e"] == 0.0) & (res["log_l"] == 0.0) & (res["var_e"] == 0.0) & (gap < 80)
        assert (gap[cert] > gthr).all(), (seed, float(gap[cert].min()))
        total_cert += eng.screen_report["rows_certified"]
        # arg-min of certified rows is the fp64 arg-min
        amin64 = e.argmin(dim=1).view(len(temp), b)
        closed = (res["l"] == 1.0) & (res["log_l"] == 0.0)
        assert torch.equal(res["argmin"][closed & (gap > 1e-3)], amin64[closed & (gap > 1e-3)]), seed
    assert total_cert > 200


def test_block_screened_in_chunks_from_its_low_temperature_end(monkeypatch):
    """A block of 16 row tiles is screened in quarters: the high-temperature quarter first; if that fails, up from the
    low-temperature end until a quarter is mostly unproven.  The full pass is one launch over listed + unscreened tiles.
    Ascending and descending order of the same schedules."""
    monkeypatch.setattr(PosteriorEngine, "SCREEN_MIN_CHUNK_TILES", 1)
    data, x0 = _setup(n=400, d=96, b=16)                        # 16 rows per temperature = 2 row tiles of 8
    b = len(x0)
    cases = (
        # schedule (quarters of 2 temperatures)                      one-product launches, unscreened rows, certified >=
        (torch.tensor([1e-4, 1e-3, 10.0, 100.0, 1e3, 1e4, 1e5, 1e6]), 3, 2 * b, 2 * b - 2),   # top fails, q0 ok, q1 fails: stop
        (torch.tensor([1e-4, 1e-3, 1e-2, 0.1, 10.0, 100.0, 1e3, 1e4]), 4, 0, 4 * b - 4),      # boundary in the third quarter
        (torch.tensor([1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 2e-2, 5e-2, 0.1]), 2, 0, 8 * b - 8),     # top proven: the rest in one launch
    )
    for asc, want_x1, want_unscreened, want_cert in cases:
        for temp in (asc, asc.flip(0)):
            eng_s, be_s, r_s, _ = _run(data, x0, temp, True, len(temp))      # ONE block
            eng_u, _, r_u, _ = _run(data, x0, temp, False, len(temp))
            rep = eng_s.screen_report
            assert be_s.calls.count("stats:f16x1:all") == want_x1, (temp, be_s.calls)
            assert be_s.calls.count("stats:f16x3:list") <= 1 and "stats:f16x3:all" not in be_s.calls
            assert rep["rows_unscreened"] == want_unscreened and rep["rows_certified"] >= want_cert, (temp, rep)
            assert rep["rows_screened"] + rep["rows_unscreened"] == len(temp) * b
            for k in ("log_l", "mean_e", "mean_e2", "var_e", "entropy", "l"):
                assert torch.isfinite(r_s[k]).all(), k
                assert torch.allclose(r_s[k], r_u[k], rtol=1e-4, atol=2e-6), (k, (r_s[k] - r_u[k]).abs().max())
            assert torch.equal(r_s["argmin"], r_u["argmin"])


def test_e4m3_cascade_is_sound_and_falls_through_to_the_fp16_stage():
    """Cascade: E4M3 stage on every row, fp16 one-product stage on the tiles it leaves, full pass on the rest.  Every
    certified row is checked in fp64; the cascade certifies at least what the fp16 stage alone certifies."""
    total8 = 0
    for seed in range(6):
        g = syn.gen(200 + seed)
        n, d, b = 350 + 31 * seed, 64 + 16 * seed, 16
        data = _dataset(seed % 4, n, d, g)
        x0 = data[torch.randint(0, n, (b,), generator=g)].clone()
        temp = torch.logspace(-5, 3, 17)
        eng8, be8, r8, noise = _run(data, x0, temp, True, 17, f8=True)
        eng1, _, r1, _ = _run(data, x0, temp, True, 17, f8=False)
        eng_u, _, r_u, _ = _run(data, x0, temp, False, 17)
        assert "stats:f8x1:all" in be8.calls
        gthr = 17.0 + math.log(n)
        xt = (noise.double() * temp.double().sqrt()[:, None, None] + x0.double()[None])
        e = 0.5 * orc.pairwise_sqdist(xt.reshape(-1, d), data.double())
        srt = e.sort(dim=1).values
        gap = ((srt[:, 1] - srt[:, 0]) / temp.double().repeat_interleave(b)).view(len(temp), b)
        cert = (r8["l"] == 1.0) & (r8["mean_e"] == 0.0) & (r8["log_l"] == 0.0) & (r8["var_e"] == 0.0) & (gap < 80)
        assert (gap[cert] > gthr).all(), (seed, float(gap[cert].min()))
        assert eng8.screen_report["rows_certified"] >= eng1.screen_report["rows_certified"], seed
        total8 += eng8.screen_report["f8_tiles_screened"] - eng8.screen_report["f8_tiles_left"]
        for k in ("log_l", "mean_e", "entropy"):
            assert torch.isfinite(r8[k]).all()
            assert torch.allclose(r8[k], r_u[k], rtol=1e-4, atol=2e-6), (seed, k)
        ok = gap > 1e-3
        assert torch.equal(r8["argmin"][ok], r_u["argmin"][ok]), seed
    assert total8 > 0                                        # the E4M3 stage proved tiles on its own


def test_posterior_mean_of_a_fully_proven_block_is_a_gather():
    """Low-noise ideal-denoiser call: every row certified -> x0_hat = the nearest training points.  The path never reads a
    count back (ABI v4: tile lists with device-side lengths), so the full-precision pass, the weights and the second
    contraction are still ENQUEUED -- over empty tile lists."""
    data, _ = _setup(n=300, d=64, b=8)
    g = syn.gen(77)
    idx = torch.tensor([1, 7, 50, 120, 299, 200, 10, 11])         # none of them the duplicated point 5 / 37
    for f8 in (False, True):
        be = SplitFakeBackend()
        eng = PosteriorEngine(EmpiricalDataset(data, backend=be), EngineConfig(precision="f16x3", screen=True, screen_f8=f8))
        ab = torch.tensor(0.9999)
        xt = ab.sqrt() * data[idx] + (1 - ab).sqrt() * torch.randn(len(idx), 64, generator=g)
        got = eng.posterior_mean(xt, ((1 - ab) / ab).expand(len(idx)), post=ab.rsqrt().expand(len(idx)))
        assert torch.equal(got, data[idx])
        want = orc.posterior_mean_x0(xt, ab, data, dtype=torch.float64)
        assert (got.double() - want).abs().max() < 1e-6
        eng._pm_poll(wait=True)                                   # the counts arrive behind the call (lagged feedback)
        assert eng.screen_report["pm_rows_certified"] == len(idx)
        assert ("stats:f8x1:all" in be.calls) == f8
        assert "tiles:f16x3:0" in be.calls and "weights:0" in be.calls and "gemm2:0" in be.calls


def test_posterior_mean_mixed_block_contracts_only_the_listed_tiles():
    """Rows of both kinds in one call: proven rows are gathered, the tile with the near-tie row goes through the
    full-precision pass, the weights and the second contraction -- and nothing else does."""
    data, _ = _setup(n=300, d=64, b=8)
    g = syn.gen(78)
    idx = torch.tensor([1, 7, 50, 120, 299, 200, 10, 11, 5, 60, 61, 62, 63, 64, 65, 66])     # row 8 sits on the duplicate 5 / 37
    be = SplitFakeBackend()
    eng = PosteriorEngine(EmpiricalDataset(data, backend=be), EngineConfig(precision="f16x3", screen=True, screen_f8=False))
    ab = torch.tensor(0.9999)
    xt = ab.sqrt() * data[idx] + (1 - ab).sqrt() * torch.randn(len(idx), 64, generator=g)
    got = eng.posterior_mean(xt, ((1 - ab) / ab).expand(len(idx)), post=ab.rsqrt().expand(len(idx)))
    want = orc.posterior_mean_x0(xt, ab, data, dtype=torch.float64)
    assert (got.double() - want).abs().max() < 1e-5
    assert torch.equal(got[:8], data[idx[:8]])                     # first tile: every row proven
    assert "tiles:f16x3:1" in be.calls and "weights:1" in be.calls and "gemm2:1" in be.calls


def test_e4m3_stage_has_a_remembered_mark_of_its_own():
    """Remembered-boundary path with the cascade on: the first such call runs the E4M3 stage on the whole certifiable range and
    learns where the stage alone stops proving rows; later calls start the rows between the two marks at the fp16 stage.  Same
    results, fewer rows through the E4M3 pass, and the stage's tile accounting follows its own span."""
    g = syn.gen(312)
    n, d, b = 400, 256, 32
    data = _dataset(2, n, d, g)
    x0 = data[torch.randint(0, n, (b,), generator=g)].clone()
    temp = torch.logspace(-5, 3, 24)
    be = SplitFakeBackend()
    cfg = EngineConfig(precision="f16x3", screen=True, screen_f8=True)
    cfg.max_query_bytes = 8 * b * d * 12                                   # blocks of eight temperatures
    eng = PosteriorEngine(EmpiricalDataset(data, backend=be), cfg)
    noise = torch.randn(len(temp), b, d, generator=syn.gen(311))
    runs = []
    for _ in range(4):
        n0 = len(be.calls)
        rep0 = dict(eng.screen_report)
        res = eng.noised_stats(x0, temp, noise_fn=lambda i: noise[i])
        runs.append((res, be.calls[n0:], eng._screen_prior, eng._screen_prior_f8,
                     eng.screen_report.get("f8_tiles_screened", 0) - rep0.get("f8_tiles_screened", 0)))
    (r1, _, p1, f1, _), (r2, c2, p2, f2, t2), (r3, c3, p3, f3, t3), (r4, c4, p4, f4, t4) = runs
    assert p1 is not None and f1 is None                                   # the probing call knows nothing about the stage
    assert f2 is not None and 0 < f2 <= p2, (f2, p2)                       # learnt by the first remembered-boundary call
    assert t3 < t2 and t4 <= t3, (t2, t3, t4)                              # fewer row tiles through the E4M3 pass afterwards
    assert any(c.startswith("stats:f8x1") for c in c3)
    plain = PosteriorEngine(EmpiricalDataset(data, backend=SplitFakeBackend()), EngineConfig(precision="f16x3", screen=False))
    ref = plain.noised_stats(x0, temp, noise_fn=lambda i: noise[i])
    for res in (r2, r3, r4):
        for k in ("log_l", "mean_e", "var_e", "entropy"):
            assert torch.isfinite(res[k]).all(), k
            assert torch.allclose(res[k], ref[k], rtol=1e-4, atol=2e-6), k
        assert torch.equal(res["argmin"], ref["argmin"])
    # a descending schedule takes the same path (the E4M3 rows sit at the END of a block)
    r5 = eng.noised_stats(x0, temp.flip(0), noise_fn=lambda i: noise[len(temp) - 1 - i])
    assert torch.allclose(r5["entropy"].flip(0), ref["entropy"], rtol=1e-4, atol=2e-6)


def test_unproven_rows_are_compacted_and_an_overflowing_buffer_is_harmless():
    """Remembered-boundary blocks gather the unproven rows of the screened range into dense row tiles for the full-precision
    pass (index list, count and tile count on the device).  The buffer is sized from the previous call's count; rows beyond it
    stay on the tile-list path -- forced here by a hint of zero rows -- and the results do not change."""
    data, x0 = _setup(n=400, d=96, b=16)
    for j in (41, 42, 43):                                   # four duplicated training points: the rows of queries 5..8 never certify
        data[j] = data[j - 35]
    b, d = x0.shape
    temp = torch.logspace(-4, 3, 36)
    noise = torch.randn(len(temp), b, d, generator=syn.gen(321))
    plain = PosteriorEngine(EmpiricalDataset(data, backend=SplitFakeBackend()), EngineConfig(precision="f16x3", screen=False))
    ref = plain.noised_stats(x0, temp, noise_fn=lambda i: noise[i])
    results = {}
    for compact in (True, False):
        be = SplitFakeBackend()
        cfg = EngineConfig(precision="f16x3", screen=True, screen_f8=False, screen_compact=compact)
        eng = PosteriorEngine(EmpiricalDataset(data, backend=be), cfg)
        eng.noised_stats(x0, temp, noise_fn=lambda i: noise[i])                  # probing call
        eng.noised_stats(x0, temp, noise_fn=lambda i: noise[i])                  # first remembered-boundary call
        rep0 = dict(eng.screen_report)
        results[compact] = eng.noised_stats(x0, temp, noise_fn=lambda i: noise[i])
        full = eng.screen_report["tiles_full_pass"] - rep0["tiles_full_pass"]
        open_rows = (eng.screen_report["rows_screened"] - rep0["rows_screened"]) - (eng.screen_report["rows_certified"] - rep0["rows_certified"])
        assert open_rows > 0
        if compact:
            assert full == -(-open_rows // be.row_tile), (full, open_rows)          # dense tiles
            assert all(isinstance(v, tuple) and v[1] == open_rows for v in eng._screen_hint.values())
            eng._screen_hint = {k: (v[0], 0) for k, v in eng._screen_hint.items()}  # next call: a buffer of two row tiles
            assert open_rows > 2 * be.row_tile
            results["overflow"] = eng.noised_stats(x0, temp, noise_fn=lambda i: noise[i])
        else:
            assert full >= -(-open_rows // be.row_tile)
    xn = ((noise * temp.sqrt()[:, None, None] + x0[None]) ** 2).sum(-1)
    for name, res in results.items():
        for k in ("log_l", "mean_e", "var_e", "entropy"):
            assert torch.isfinite(res[k]).all(), (name, k)
            assert torch.allclose(res[k], ref[k], rtol=1e-4, atol=2e-6), (name, k)
        assert ((res["e_min"] - ref["e_min"]).abs() <= 8 * 2.0 ** -24 * (xn + d)).all(), name     # proven rows: E_min recomputed in fp64
        assert torch.equal(res["argmin"], ref["argmin"]), name
    for k in ("log_l", "mean_e", "var_e", "entropy", "e_min"):                     # same arithmetic row by row: identical
        assert torch.equal(results[True][k], results[False][k]), k
        assert torch.equal(results["overflow"][k], results[False][k]), k
