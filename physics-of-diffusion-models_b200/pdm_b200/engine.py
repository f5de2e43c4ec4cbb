"""Host-side orchestration of the empirical-denoiser hot path.

``EmpiricalDataset`` keeps one shard of the training set resident in HBM together with everything that
is computed once per dataset (row norms, the global power-of-two scale, the fp16 hi/lo operand split, the
transposed split for the posterior-mean contraction, Tr Sigma_0).  ``PosteriorEngine`` runs the fused
pass for a block of query rows -- all temperatures of a schedule are flattened into the rows of one
launch -- merges the per-split / per-shard partial records and finalises the quantities the reference
reports.  Every numerical step happens in the CUDA library (``backend``); torch provides memory, streams,
the RNG stream of the reference (``torch.randn`` per temperature) and ``torch.distributed`` plumbing.

Reference call sites this replaces: utils/stats.py:71-111, 271-290 (per-temperature loops),
diffusion/scheduler/scheduler.py:60-69 (ideal denoiser), utils/distance.py:13-21.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Optional

import torch
from torch import Tensor

from . import _cabi
from ._cabi import PdmError

STAT_KEYS = ("e_min", "log_l", "mean_e", "mean_e2", "var_e", "aux_mean", "entropy", "l")


def default_backend():
    from .backend import CudaBackend
    return CudaBackend()


def _env_int(name: str, default: int = 0) -> int:
    v = os.environ.get(name)
    return int(v) if v else default


def _flat2d(t: Tensor) -> Tensor:
    """(rows, everything else) -- also for tensors without rows, where reshape(0, -1) is ambiguous."""
    return t.reshape(t.shape[0], math.prod(t.shape[1:]))


def pow2_scale_for(absmax: float) -> float:
    """2^k with absmax * 2^k in [2^11, 2^12): the fp16 hi part keeps 11 bits, hi+lo 22 bits."""
    if not (absmax > 0.0) or math.isinf(absmax) or math.isnan(absmax):
        return 1.0
    _, e = math.frexp(absmax)           # absmax = f * 2^e, f in [0.5, 1)
    return math.ldexp(1.0, max(-100, min(100, 12 - e)))


# ||y_j - hi_j/s|| <= 2^-22 ||y_j|| for every row: the dropped part is a few units of fp32 round-off of the row
# itself (for pixel data hi/s IS the exact pixel value and the residue is the fp32 rounding of the transform);
# worst case it moves an energy by 2^-23 (||x||^2 + ||y||^2), 1/4 of the round-off floor of the parity contract.
LATTICE_RATIO_MAX = 2.0 ** -44
LATTICE_INT_MAX = 2048.0           # fp16 holds integers up to 2^11 exactly


def lattice_candidates(absmax: float) -> list[float]:
    """Scales s for which y*s could be an fp16-exact integer lattice: the 8-bit pixel grid behind
    ToTensor / Normalize(0.5, 0.5) (utils/data.py:43-52 of the reference: s = 255 * 2^k) and dyadic data (s = 2^k),
    largest first, restricted to absmax * s <= 2048."""
    if not (absmax > 0.0) or math.isinf(absmax) or math.isnan(absmax):
        return []
    out = []
    for base in (1.0, 255.0):
        k = math.floor(math.log2(LATTICE_INT_MAX / (absmax * base)))
        for kk in range(k, k - 4, -1):
            s = base * 2.0 ** kk
            if absmax * s <= LATTICE_INT_MAX and 2.0 ** -100 < s < 2.0 ** 100:
                out.append(s)
    return out


def detect_lattice_scale(backend, y: Tensor, absmax: Optional[float] = None) -> float:
    """The scale under which the rows of y are an fp16-exact lattice, or 0.0."""
    if absmax is None:
        absmax = float(backend.absmax(y).item())
    for s in lattice_candidates(absmax):
        ratio, vmax = backend.lattice_residual(y, s).tolist()
        if ratio <= LATTICE_RATIO_MAX and vmax <= LATTICE_INT_MAX:
            return s
    return 0.0


class EmpiricalDataset:
    """A (shard of a) training set resident on one GPU."""

    def __init__(self, data: Tensor, *, backend=None, index_offset: int = 0, n_total: Optional[int] = None,
                 global_absmax: Optional[float] = None, lattice_scale: Optional[float] = None):
        """``global_absmax`` / ``lattice_scale``: whole-dataset facts when ``data`` is one shard of a row-sharded
        set (every rank must use the same scale; lattice_scale 0.0 = not a lattice, None = detect on ``data``)."""
        self.backend = backend if backend is not None else default_backend()
        dev = self.backend.device
        flat = _flat2d(data)
        self.y = flat.to(device=dev, dtype=torch.float32).contiguous()
        self.item_shape = tuple(data.shape[1:])
        self.n, self.d = self.y.shape
        self.index_offset = int(index_offset)
        self.n_total = int(n_total) if n_total is not None else self.n
        self.y_norm = self.backend.row_norms(self.y)
        self._global_absmax = global_absmax
        self._lattice = lattice_scale
        self._split = None
        self._tsplit = None
        self._yt = None
        self._moments = None
        self._scale = None
        self._y_norm_max = None
        self._e4m3 = None

    def e4m3(self):
        """(E4M3 bytes of the dataset split, max_j of the rows' exact rounding deviation in data units, one-element
        device tensor): operands of the first stage of the screening cascade."""
        if self._e4m3 is None:
            y_hi, y_lo = self.split()
            y8, err = self.backend.split_to_e4m3(y_hi, y_lo, self.d)
            self._e4m3 = (y8, (err.max() / self.scale).reshape(1).contiguous())
        return self._e4m3

    @property
    def y_norm_max(self) -> Tensor:
        """max_j ||y_j||^2 as a one-element device tensor (error bound of the screening pass)."""
        if self._y_norm_max is None:
            self._y_norm_max = self.y_norm.max().reshape(1).contiguous()
        return self._y_norm_max

    # -- lazily built device-side views ----------------------------------------------------------
    def _absmax(self) -> float:
        if self._global_absmax is None:
            self._global_absmax = float(self.backend.absmax(self.y).item())
        return self._global_absmax

    @property
    def lattice_scale(self) -> float:
        """s > 0 when y*s is an fp16-exact integer lattice (8-bit image data): the lo part of the split is
        empty and the two-product mode f16x2 loses nothing.  0.0 otherwise."""
        if self._lattice is None:
            detect = getattr(self.backend, "lattice_residual", None)
            self._lattice = detect_lattice_scale(self.backend, self.y, self._absmax()) if detect is not None else 0.0
        return self._lattice

    @property
    def scale(self) -> float:
        if self._scale is None:
            self._scale = self.lattice_scale or pow2_scale_for(self._absmax())
        return self._scale

    def split(self):
        """(y_hi, y_lo) fp16 operands of y * scale, shape (n, round_up(d, 8))."""
        if self._split is None:
            r = self.backend.prepare_rows(self.y, self.n, fixed_scale=self.scale, want_norms=False)
            self._split = (r["hi"], r["lo"])
        return self._split

    def transposed_split(self):
        if self._tsplit is None:
            self._tsplit = self.backend.transpose_split(self.y, self.scale)
        return self._tsplit

    def transposed(self) -> Tensor:
        """y^T (d, n) fp32, for the exact-path contraction g . Y^T of the denoiser's backward pass."""
        if self._yt is None:
            self._yt = self.y.t().contiguous()
        return self._yt

    def moments(self):
        """(column sums, column sums of squares, [min, max]) -- fp64 / fp32 device tensors."""
        if self._moments is None:
            self._moments = self.backend.column_moments(self.y)
        return self._moments

    def tr_sigma0(self) -> float:
        """sum_k Var_k (unbiased), torch.var(x, dim=0).sum() of utils/stats.py:39,176 (single-shard form)."""
        s, s2, _ = self.moments()
        n = float(self.n)
        return float(((s2 - s * s / n) / (n - 1.0)).sum().item())

    def value_range(self):
        mm = self.moments()[2]
        return float(mm[0].item()), float(mm[1].item())


@dataclass
class EngineConfig:
    precision: str = "auto"            # auto | exact | f16x3 | f16x2 (lattice datasets only) | f16x1
    tensor_min_dim: int = 64           # auto: below one 64-wide k-block the exact CUDA-core kernel is used (at
                                       # d = 64..192 the tensor path measured 4-6x faster AND closer to fp64)
    cta_group: int = 0                 # 0 = library default (2)
    m_group: int = 0
    n_splits: int = 0
    max_query_bytes: int = 6 << 30     # scratch budget for one block of query rows (noise + split): fewer, larger
                                       # launches (and fewer cross-GPU merges) per schedule; 180 GB of HBM3e to spare
    max_energy_bytes: int = 8 << 30    # scratch budget for the energy tile + weights of the posterior-mean path: 10^4 queries against
                                       # N = 50 000 in ONE block (two blocks of 8053 + 1947 rows cost a whole extra round of tiles)
    sync_noise: bool = True            # sharded runs: make every rank draw rank 0's noise stream (set False when
                                       # every rank seeds its generator identically)
    delta_shortcut: bool = True        # posterior mean: rows whose posterior is a delta to fp32 resolution take their
                                       # nearest training point directly (skips weights + second contraction for them)
    fused_noise: bool = True           # regenerate torch.randn's Philox stream inside the operand kernel (bit-identical,
                                       # verified once per device) instead of torch.randn + pdm_prepare_rows
    screen: bool = True                # certified delta posteriors (on by default since round 2: parity at the named
                                       # configurations is pinned with it on, tests/test_gpu_named_configs.py).  A one-product
                                       # tensor pass with a rigorous error bound proves, row by row, that every other training
                                       # point's weight is below exp(-screen_g) of the nearest one's; proven rows take the closed
                                       # form and only the row tiles with an unproven row run the full-precision pass
                                       # (include/pdm_b200.h, pdm_screen_*).  PDM_SCREEN=0 turns it off.
    screen_f8: bool = True             # screening cascade: an E4M3 pass (twice the MMA rate) first, the fp16 one-product pass
                                       # only on the row tiles it leaves unproven (PDM_SCREEN_F8=0 turns the first stage off)
    screen_g: float = 0.0              # weight cut-off exponent; 0 = 17 + log N (everything dropped sums to < 2^-24)
    screen_compact: bool = True        # remembered-boundary blocks: the unproven rows of the screened range are gathered into
                                       # dense row tiles before the full-precision pass (near the boundary a 256-row tile
                                       # holds proven and unproven rows side by side).  PDM_SCREEN_COMPACT=0 turns it off.

    @staticmethod
    def from_env() -> "EngineConfig":
        cfg = EngineConfig(precision=os.environ.get("PDM_PRECISION", "auto"),
                           cta_group=_env_int("PDM_CTA_GROUP"), m_group=_env_int("PDM_M_GROUP"),
                           n_splits=_env_int("PDM_N_SPLITS"))
        if os.environ.get("PDM_SCREEN", "") in ("0", "1"):
            cfg.screen = os.environ["PDM_SCREEN"] == "1"
        if os.environ.get("PDM_SCREEN_F8", "") in ("0", "1"):
            cfg.screen_f8 = os.environ["PDM_SCREEN_F8"] == "1"
        if os.environ.get("PDM_SCREEN_COMPACT", "") in ("0", "1"):
            cfg.screen_compact = os.environ["PDM_SCREEN_COMPACT"] == "1"
        return cfg


class PosteriorEngine:
    """Fused posterior statistics / posterior mean of query rows against one dataset shard.

    ``group``: optional torch.distributed process group over which the dataset is row-sharded; every rank of it
    must present the same query rows, partial records are all-gathered and merged.
    ``query_group``: optional group of ranks that hold the SAME dataset shard and split the work of ``noised_stats``
    between them: the temperatures of the schedule are dealt round-robin, each rank draws and evaluates only its own
    (at the Philox offsets they have in the one stream a single-GPU run consumes) and the (n_T, B) results are
    all-gathered at the end of the call (pdm_b200/sharding.py builds both groups).
    """

    # test hook: callable (temperature index, shape, device) -> standard-normal tensor replacing torch.randn
    noise_hook = None

    def __init__(self, dataset: EmpiricalDataset, config: Optional[EngineConfig] = None, group=None, query_group=None):
        self.ds = dataset
        self.backend = dataset.backend
        self.cfg = config if config is not None else EngineConfig.from_env()
        self.group = group
        self.world = 1
        self.query_group = query_group
        self.q_world, self.q_rank = 1, 0
        # screening bookkeeping: rows / row tiles seen by the screening pass and what it left for the full pass
        self.screen_report = {"rows_screened": 0, "rows_certified": 0, "tiles_screened": 0, "tiles_full_pass": 0,
                              "rows_unscreened": 0}
        self._screen_t_fail = math.inf
        self._y_err8_max = None
        self._screen_f8_live = True        # the E4M3 first stage is still proving enough in this call
        self._pm_f8_t = math.inf           # posterior_mean: blocks below this mark try the E4M3 stage first
        self._pm_screen_t = math.inf       # posterior_mean: blocks whose temperatures are all below this mark are screened
        self._screen_t_retry = math.inf    # after a screened block that certified nothing: next attempt at T <= this
        self._pm_pending: list = []        # posterior_mean: screening counts in flight to the host (lagged feedback)
        self._screen_prior = None          # noised_stats: lowest temperature at which a row stayed unproven (across calls)
        self._screen_prior_f8 = None       # noised_stats: the same mark for the E4M3 first stage alone (rows between the two marks skip it)
        self._screen_hint: dict = {}       # noised_stats: listed tiles of a block in the previous call (schedule hint)
        self._y_norm_max = None
        if group is not None:
            import torch.distributed as dist
            self.world = dist.get_world_size(group)
        if query_group is not None:
            import torch.distributed as dist
            self.q_world, self.q_rank = dist.get_world_size(query_group), dist.get_rank(query_group)

    def _local_aux(self, aux: Optional[Tensor]) -> Optional[Tensor]:
        """The per-point aux vector as the kernels index it: by LOCAL dataset row.  A vector over the whole (row-sharded)
        dataset is cut to this rank's rows (a fresh tensor: the kernels need 16-byte alignment)."""
        if aux is None:
            return None
        ds = self.ds
        aux = aux.to(device=self.backend.device, dtype=torch.float32).reshape(-1)
        if aux.numel() == ds.n_total and ds.n_total != ds.n:
            return aux[ds.index_offset:ds.index_offset + ds.n].clone()
        if aux.numel() != ds.n:
            raise PdmError(f"aux must hold one value per dataset row: got {aux.numel()}, this shard has {ds.n} rows "
                           f"of {ds.n_total}")
        return aux.contiguous()

    # -- precision -------------------------------------------------------------------------------
    def precision(self) -> str:
        p = self.cfg.precision
        if p == "auto":
            tensor_ok = getattr(self.backend, "supports_tensor_path", lambda: False)()
            if not (self.ds.d >= self.cfg.tensor_min_dim and tensor_ok):
                return "exact"
            return "f16x2" if self.ds.lattice_scale > 0 else "f16x3"
        if p not in ("exact", "f16x3", "f16x2", "f16x1"):
            raise PdmError(f"unknown precision {p!r}")
        if p == "f16x2" and not self.ds.lattice_scale > 0:
            raise PdmError("precision f16x2 drops the lo part of the dataset split: it needs a dataset that is an "
                           "fp16-exact lattice (EmpiricalDataset.lattice_scale > 0), e.g. 8-bit images")
        return p

    def rows_per_block(self, row_multiple: int = 1) -> int:
        per_row = self.ds.d * 4 * 3
        rows = max(1, self.cfg.max_query_bytes // per_row)
        return max(row_multiple, rows // row_multiple * row_multiple)

    # -- core: one block of query rows -----------------------------------------------------------
    def _prepare(self, src: Tensor, rows: int, noise, sigma, post, precision: str, want_x: bool):
        tensor = precision != "exact"
        return self.backend.prepare_rows(src, rows, noise=noise, sigma=sigma, post=post,
                                         want_x=(want_x or not tensor), want_norms=True, want_split=tensor)

    def _local_partials(self, prep: dict, rows: int, inv_temp: Tensor, aux: Optional[Tensor], precision: str,
                        energy_out: Optional[Tensor] = None, energy_mult: float = 1.0, want_partials: bool = True,
                        row_tiles: Optional[Tensor] = None, n_row_tiles: int = 0, n_row_tiles_dev: Optional[Tensor] = None,
                        plan_row_tiles: int = 0):
        ds = self.ds
        kw = dict(precision=precision, M=rows, N=ds.n, d=ds.d, q_norm=prep["norms"], y_norm=ds.y_norm,
                  inv_temp=inv_temp, y_aux=aux, index_offset=ds.index_offset, n_splits=self.cfg.n_splits,
                  m_group=self.cfg.m_group, cta_group=self.cfg.cta_group, want_partials=want_partials,
                  energy_out=energy_out, energy_mult=energy_mult)
        if row_tiles is not None:
            kw.update(row_tiles=row_tiles, n_row_tiles=n_row_tiles)
            if n_row_tiles_dev is not None:
                kw.update(n_row_tiles_dev=n_row_tiles_dev)
            if plan_row_tiles > 0:
                kw.update(plan_row_tiles=plan_row_tiles)
        if precision == "exact":
            return self.backend.posterior_stats(q=prep["x"], y=ds.y, **kw)
        y_hi, y_lo = ds.split()
        return self.backend.posterior_stats(q_split=(prep["hi"], prep["lo"], prep["inv_scale"]),
                                            y_split=(y_hi, None if precision == "f16x2" else y_lo),
                                            y_inv_scale=1.0 / ds.scale, **kw)

    def _merge(self, parts: Tensor, inv_temp: Tensor):
        if self.world > 1:
            import torch.distributed as dist
            local = self.backend.reduce(parts, inv_temp)          # 32 B per row cross the link, not 32 B per split
            gathered = torch.empty((self.world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype,
                                   device=local.device)
            dist.all_gather_into_tensor(gathered, local, group=self.group)     # rank-major concatenation
            parts = gathered.view((self.world,) + tuple(local.shape))
        return self.backend.merge(parts, inv_temp, self.ds.n_total)

    # -- certified delta posteriors ----------------------------------------------------------------
    SCREEN_E_STAR = 50.0       # the screening pass certifies (E1_j - E1_min)/T' > e_star for every other point
    SCREEN_KAPPA = 1.25        # safety factor on the first-order bound 2^-10 ||x|| ||y|| of |x.y - x_hi.y_hi|
    SCREEN_MIN_CHUNK_TILES = 32   # a block is screened in four chunks when each holds at least this many row tiles

    def screening_usable(self) -> bool:
        return (self.cfg.screen and self.precision() in ("f16x3", "f16x2") and hasattr(self.backend, "screen_certify"))

    def _global_y_norm_max(self) -> Tensor:
        if self._y_norm_max is None:
            v = self.ds.y_norm_max.clone()
            if self.world > 1:
                import torch.distributed as dist
                dist.all_reduce(v, op=dist.ReduceOp.MAX, group=self.group)
            self._y_norm_max = v
        return self._y_norm_max

    SCREEN_KAPPA_F8 = 1.02     # the E4M3 stage's bound uses the exact rounding deviations; this covers the accumulation

    def _global_max(self, v: Tensor) -> Tensor:
        v = v.clone()
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(v, op=dist.ReduceOp.MAX, group=self.group)
        return v

    def _screen_certificate(self, prep: dict, rows: int, inv_temp: Tensor, stage: str = "f16x1",
                            row_tiles: Optional[Tensor] = None, n_row_tiles: int = 0, n_row_tiles_dev: Optional[Tensor] = None):
        """One-product pass (stage "f16x1": fp16 hi parts; "f8x1": E4M3 bytes) at the fictitious temperature + certificate:
        (flags, arg-min of the pass, list of row tiles with an unproven row, its length (device), rows per tile).
        With ``row_tiles`` the pass covers the listed row tiles only; flags of other rows are meaningless."""
        be, ds = self.backend, self.ds
        g = self.cfg.screen_g if self.cfg.screen_g > 0 else 17.0 + math.log(max(2, ds.n_total))
        kw = {} if row_tiles is None else dict(row_tiles=row_tiles, n_row_tiles=n_row_tiles)
        if row_tiles is not None and n_row_tiles_dev is not None:
            kw["n_row_tiles_dev"] = n_row_tiles_dev
        if stage == "f8x1":
            if self._y_err8_max is None:
                self._y_err8_max = self._global_max(ds.e4m3()[1])
            q8, q_err = be.split_to_e4m3(prep["hi"], prep["lo"], ds.d)
            inv_t1 = be.screen_temperatures_f8(prep["norms"], q_err, prep["inv_scale"], inv_temp, self._global_y_norm_max(),
                                               self._y_err8_max, g, self.SCREEN_E_STAR, self.SCREEN_KAPPA_F8)
            parts1 = be.posterior_stats(precision="f8x1", M=rows, N=ds.n, d=ds.d, q_norm=prep["norms"], y_norm=ds.y_norm,
                                        inv_temp=inv_t1, q_split=(q8, None, prep["inv_scale"] * 16.0),
                                        y_split=(ds.e4m3()[0], None), y_inv_scale=16.0 / ds.scale,
                                        index_offset=ds.index_offset, cta_group=self.cfg.cta_group, **kw)
        else:
            inv_t1 = be.screen_temperatures(prep["norms"], inv_temp, self._global_y_norm_max(), g, self.SCREEN_E_STAR,
                                            self.SCREEN_KAPPA)
            parts1 = self._local_partials(prep, rows, inv_t1, None, "f16x1", **kw)
        out1, arg1 = self._merge(parts1, inv_t1)             # across shards too: every rank sees the same certificate
        rows_per_tile = getattr(be, "row_tile", None) or 128 * (self.cfg.cta_group or 2)     # the fused kernel's row tile
        flags, tile_list, n_listed = be.screen_certify(out1, self.SCREEN_E_STAR, rows_per_tile)
        return flags, arg1, tile_list, n_listed, rows_per_tile

    def _f8_stage_usable(self) -> bool:
        return self.cfg.screen_f8 and hasattr(self.backend, "split_to_e4m3") and self.ds.d % 8 == 0

    def _screen_cascade(self, prep: dict, rows: int, inv_temp: Tensor, use_f8: bool):
        """E4M3 stage on every row (``use_f8``), fp16 one-product stage on the row tiles it leaves unproven.  Returns the
        values of _screen_certificate plus the fraction of row tiles the E4M3 stage left unproven (None: stage not run)."""
        be = self.backend
        if not (use_f8 and self._f8_stage_usable()):
            return self._screen_certificate(prep, rows, inv_temp) + (None,)
        f8, a8, tl8, nl8, rpt = self._screen_certificate(prep, rows, inv_temp, stage="f8x1")
        n8 = int(nl8.item())
        tiles = (rows + rpt - 1) // rpt
        rep = self.screen_report
        rep["f8_tiles_screened"] = rep.get("f8_tiles_screened", 0) + tiles
        rep["f8_tiles_left"] = rep.get("f8_tiles_left", 0) + n8
        if n8 == 0:
            return f8, a8, tl8, nl8, rpt, 0.0
        f1, a1, _, _, _ = self._screen_certificate(prep, rows, inv_temp, row_tiles=tl8, n_row_tiles=n8)
        in_list = torch.zeros(tiles, dtype=torch.bool, device=be.device)
        in_list[tl8[:n8].long()] = True
        row_in_list = in_list.repeat_interleave(rpt)[:rows]
        proven8 = f8 != 0
        flags = (proven8 | (row_in_list & (f1 != 0))).to(torch.uint8)
        arg = torch.where(proven8, a8, a1)
        tile_list, n_listed = be.screen_tile_list(flags, rpt)
        return flags, arg, tile_list, n_listed, rpt, n8 / tiles

    def _screen_cascade_async(self, prep: dict, rows: int, inv_temp: Tensor, use_f8: bool):
        """The cascade without a single host read (the ideal-denoiser path inside a sampling loop): the E4M3 stage's list of
        unproven row tiles and its length stay on the device and drive the fp16 one-product stage, whose verdicts are folded
        into the flags in place.  Returns (flags, arg-min of the passes, tile list, its length (device), rows per tile,
        length of the E4M3 stage's list (device) or None)."""
        be = self.backend
        if not (use_f8 and self._f8_stage_usable()):
            return self._screen_certificate(prep, rows, inv_temp) + (None,)
        f8, a8, tl8, nl8, rpt = self._screen_certificate(prep, rows, inv_temp, stage="f8x1")
        self._last_open8 = f8 == 0           # what the E4M3 stage alone left unproven (feeds its own mark in noised_stats)
        tiles = (rows + rpt - 1) // rpt
        f1, a1, _, _, _ = self._screen_certificate(prep, rows, inv_temp, row_tiles=tl8, n_row_tiles=tiles, n_row_tiles_dev=nl8)
        be.screen_merge_stage(tl8, nl8, tiles, rpt, f1, a1, f8, a8)
        tile_list, n_listed = be.screen_tile_list(f8, rpt)
        return f8, a8, tile_list, n_listed, rpt, nl8

    # lagged feedback for the screening marks of posterior_mean: the counts of a call are copied to pinned memory behind an
    # event and read when a LATER call finds the event complete -- the loop never waits for the device.
    def _pm_report(self, n_listed: Tensor, n_left8: Optional[Tensor], tiles: int, rows: int, t_lo: float, t_hi: float) -> None:
        dev = self.backend.device
        vals = torch.cat([n_listed.reshape(1), n_left8.reshape(1) if n_left8 is not None else n_listed.new_full((1,), -1)])
        if dev.type == "cuda":
            buf = torch.empty(2, dtype=torch.int32, pin_memory=True)
            buf.copy_(vals, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        else:
            buf, ev = vals.clone(), None
        self._pm_pending.append((ev, buf, tiles, rows, t_lo, t_hi))

    def _pm_poll(self, wait: bool = False) -> None:
        rep = self.screen_report
        while self._pm_pending and (wait or self._pm_pending[0][0] is None or self._pm_pending[0][0].query()):
            ev, buf, tiles, rows, t_lo, t_hi = self._pm_pending.pop(0)
            if ev is not None and wait:
                ev.synchronize()
            n_left, n8 = int(buf[0]), int(buf[1])
            rep["pm_rows_screened"] = rep.get("pm_rows_screened", 0) + rows
            rep["pm_tiles_screened"] = rep.get("pm_tiles_screened", 0) + tiles
            rep["pm_tiles_full_pass"] = rep.get("pm_tiles_full_pass", 0) + n_left
            if n_left == 0:
                rep["pm_rows_certified"] = rep.get("pm_rows_certified", 0) + rows
            if n8 >= 0:                                        # same kind of mark for the E4M3 first stage
                if n8 > 0.75 * tiles:
                    self._pm_f8_t = min(self._pm_f8_t, 0.5 * t_lo)
                elif math.isfinite(self._pm_f8_t):
                    self._pm_f8_t = max(self._pm_f8_t, 1.5 * t_hi)
            if 2 * n_left <= tiles:                            # at least half of the tiles proven: worth it up to 1.5 T
                if math.isfinite(self._pm_screen_t):
                    self._pm_screen_t = max(self._pm_screen_t, 1.5 * t_hi)
            else:                                              # a failed attempt at temperature T is repeated below T/2
                self._pm_screen_t = min(self._pm_screen_t, 0.5 * t_lo)

    @staticmethod
    def _open_runs(open_rows: Tensor, t_rows: Tensor, per_temp: int):
        """(fraction of unproven rows, temperature) per run of ``per_temp`` rows (one temperature of a schedule; 1 = every row
        on its own; a ragged tail joins nothing and is dropped)."""
        n = open_rows.numel()
        if per_temp <= 1 or n < per_temp:
            return open_rows.to(torch.float32), t_rows
        runs = n // per_temp
        frac = open_rows[:runs * per_temp].view(runs, per_temp).to(torch.float32).mean(dim=1)
        return frac, t_rows[:runs * per_temp].view(runs, per_temp)[:, 0]

    @staticmethod
    def _open_boundary(open_rows: Tensor, t_rows: Tensor, per_temp: int) -> Tensor:
        """Lowest temperature at which MOST rows stayed unproven (device scalar, +inf if none): the boundary of the
        certifiable range.  Single rows that can never be certified -- a query next to duplicated or near-duplicate
        training points fails at every temperature -- must not move it, so a temperature counts as failed only when more
        than half of its rows are open."""
        frac, t_run = PosteriorEngine._open_runs(open_rows, t_rows, per_temp)
        return torch.where(frac > 0.5, t_run, torch.full_like(t_run, math.inf)).min()

    def _screened_block(self, prep: dict, rows: int, temp_rows: Tensor, inv_temp: Tensor, aux: Optional[Tensor],
                        precision: str, ascending: bool = True, wide: bool = True, per_temp: int = 1):
        """One-product pass at the fictitious temperature -> certificate per row -> full-precision pass over the row
        tiles that hold an unproven row -> closed form for the proven rows.

        A large block that spans a wide temperature range (``wide``: more than 8x -- the wider the range, the likelier the
        certification boundary lies inside) is screened in tile-aligned quarters: the high-temperature quarter first (mostly proven -> the rest
        in one launch), otherwise up from the low-temperature end until a quarter leaves more than half of its tiles
        unproven -- a block that straddles the certification boundary pays the one-product pass for a quarter or two of
        its rows beyond the boundary, not for all of them.  The full pass is ONE launch over the listed tiles of the
        screened quarters plus every tile of the unscreened ones.  One host read (three scalars) per screening launch."""
        be, ds = self.backend, self.ds
        dev = be.device
        rpt = getattr(be, "row_tile", None) or 128 * (self.cfg.cta_group or 2)          # the fused kernel's row tile
        tiles = (rows + rpt - 1) // rpt
        n_chunks = 4 if wide and tiles >= 4 * self.SCREEN_MIN_CHUNK_TILES else 1
        bounds = [tiles * i // n_chunks for i in range(n_chunks + 1)]
        flags = torch.zeros(rows, dtype=torch.uint8, device=dev)
        arg1 = torch.zeros(rows, dtype=torch.int64, device=dev)
        listed, n_tiles = [], 0
        rep = self.screen_report

        def leave(ta, tb):                                  # tiles ta..tb-1 go to the full pass unscreened
            nonlocal n_tiles
            listed.append(torch.arange(ta, tb, dtype=torch.int32, device=dev))
            n_tiles += tb - ta
            rep["rows_unscreened"] += min(rows, tb * rpt) - ta * rpt

        def screen(ta, tb) -> int:                          # one-product pass + certificate on tiles ta..tb-1
            nonlocal n_tiles
            r0, r1 = ta * rpt, min(rows, tb * rpt)
            sub = {k: (v[r0:r1] if isinstance(v, Tensor) else v) for k, v in prep.items()}
            f, a, tl, nl, _, left8 = self._screen_cascade(sub, r1 - r0, inv_temp[r0:r1], self._screen_f8_live)
            if left8 is not None and left8 > 0.75:
                self._screen_f8_live = False                 # the E4M3 stage no longer pays in this call
            flags[r0:r1] = f
            arg1[r0:r1] = a
            open_rows = f == 0
            # rows r0.. start on a temperature boundary whenever the tile size divides per_temp (B = 1024 queries, 256-row
            # tiles); otherwise the runs are shifted by a few rows, which a majority vote does not mind
            t_open = self._open_boundary(open_rows, temp_rows[r0:r1], per_temp)
            n_l, t_open, n_open = (float(v) for v in torch.stack(
                [nl[0].to(torch.float64), t_open.to(torch.float64), open_rows.sum().to(torch.float64)]).cpu())
            n_l, n_open = int(n_l), int(n_open)
            if n_l > 0:
                listed.append(tl[:n_l] + ta)
                n_tiles += n_l
            rep["rows_screened"] += r1 - r0
            rep["rows_certified"] += (r1 - r0) - n_open
            rep["tiles_screened"] += tb - ta
            rep["tiles_full_pass"] += n_l
            self._screen_t_fail = min(self._screen_t_fail, t_open)
            if n_open == r1 - r0 and math.isfinite(t_open):
                self._screen_t_retry = min(self._screen_t_retry, 0.25 * t_open)
            elif 2 * n_l <= tb - ta:
                self._screen_t_retry = math.inf            # the certifiable range has been reached
            return n_l

        if n_chunks == 1:
            screen(0, tiles)
        else:
            # Probe the high-temperature chunk first: if at least half of its tiles are proven, so will nearly all of the rest be,
            # and the rest goes through in one launch.  Otherwise walk the remaining chunks up from the low-temperature
            # end and stop after the first one that leaves more than half of its tiles unproven.
            top = n_chunks - 1 if ascending else 0
            rest = (0, bounds[top]) if ascending else (bounds[1], tiles)
            if 2 * screen(bounds[top], bounds[top + 1]) <= bounds[top + 1] - bounds[top]:
                screen(*rest)
            else:
                stopped = False
                for c in (range(0, top) if ascending else range(n_chunks - 1, 0, -1)):
                    ta, tb = bounds[c], bounds[c + 1]
                    if stopped:
                        leave(ta, tb)
                    else:
                        stopped = 2 * screen(ta, tb) > tb - ta
        tile_list = torch.cat(listed).contiguous() if len(listed) > 1 else (listed[0].contiguous() if listed else None)
        if n_tiles > 0:
            parts = self._local_partials(prep, rows, inv_temp, aux, precision, row_tiles=tile_list, n_row_tiles=n_tiles)
            out, argmin = self._merge(parts, inv_temp)
        else:
            out = torch.empty(len(STAT_KEYS), rows, dtype=torch.float32, device=be.device)
            argmin = torch.empty(rows, dtype=torch.int64, device=be.device)
        y_hi, y_lo = ds.split()
        be.screen_finalize(flags, arg1, ds.d, (prep["hi"], prep["lo"], prep["inv_scale"]), prep["norms"],
                           (y_hi, None if precision == "f16x2" else y_lo), 1.0 / ds.scale, ds.y_norm, aux,
                           ds.index_offset, ds.n, ds.n_total, out, argmin)
        if self.world > 1:
            # the owner of a certified row's nearest point holds its E_min / aux value (others: +inf / -inf)
            import torch.distributed as dist
            both = torch.stack([out[_cabi.OUT_E_MIN], -out[_cabi.OUT_AUX_MEAN]])
            dist.all_reduce(both, op=dist.ReduceOp.MIN, group=self.group)
            out[_cabi.OUT_E_MIN].copy_(both[0])
            out[_cabi.OUT_AUX_MEAN].copy_(-both[1])
        return out, argmin

    def _screened_block_prior(self, prep: dict, rows: int, temp_rows: Tensor, inv_temp: Tensor, aux: Optional[Tensor],
                              precision: str, screen_rows: tuple, plan_tiles: int = 0, per_temp: int = 1,
                              f8_rows: Optional[tuple] = None):
        """The same pipeline once the certifiable temperature range is known from an earlier call on this dataset: rows
        ``screen_rows`` = [r0, r1) (tile-aligned, the block's low-temperature end) go through the cascade in one go -- those
        of ``f8_rows`` from the E4M3 stage on, the others from the fp16 stage -- the rest is left to the full pass, and NOTHING
        is read back: tile lists, the index list of the unproven rows and all their lengths stay on the device
        (pdm_stats_args.n_row_tiles_dev).  ``plan_tiles`` = (listed tiles, unproven rows) of the same block in the previous
        call: schedule and buffer-size hints only.  Returns (out, argmin, feedback) with feedback = device scalars (listed
        tiles, lowest mostly-unproven temperature, unproven rows, tiles the E4M3 stage left, the top screened temperature and
        how it fared, the same three numbers for the E4M3 stage alone) that the caller reads once at the end of the call."""
        be, ds = self.backend, self.ds
        dev = be.device
        rpt = getattr(be, "row_tile", None) or 128 * (self.cfg.cta_group or 2)
        tiles = (rows + rpt - 1) // rpt
        r0, r1 = screen_rows
        flags = torch.zeros(rows, dtype=torch.uint8, device=dev)          # unscreened rows count as unproven: listed
        arg1 = torch.zeros(rows, dtype=torch.int64, device=dev)
        ph = getattr(be, "phase", None)
        if ph is None:
            import contextlib
            ph = lambda _n: contextlib.nullcontext()      # noqa: E731
        # ``f8_rows`` = [a, b) inside the span: the rows below the E4M3 stage's own mark.  They go through the whole cascade;
        # the rest of the span -- temperatures the E4M3 stage is known to leave unproven -- starts at the fp16 stage.
        a8, b8 = f8_rows if (f8_rows is not None and self._screen_f8_live) else ((r0, r1) if self._screen_f8_live else (r0, r0))
        nl8 = None
        f8_fb = None                               # (lowest temperature the E4M3 stage mostly failed, its top temperature, how that fared)
        with ph("screen: cascade"):
            for (ca, cb), use8 in (((a8, b8), True), ((r0, a8), False), ((b8, r1), False)):
                if cb <= ca:
                    continue
                sub = {k: (v[ca:cb] if isinstance(v, Tensor) else v) for k, v in prep.items()}
                self._last_open8 = None
                f, a, _, _, _, n8 = self._screen_cascade_async(sub, cb - ca, inv_temp[ca:cb], use8)
                flags[ca:cb] = f
                arg1[ca:cb] = a
                if use8 and n8 is not None:
                    nl8 = n8
                    if self._last_open8 is not None:
                        fr8, tr8 = self._open_runs(self._last_open8, temp_rows[ca:cb], per_temp)
                        top8 = tr8.argmax().reshape(1)
                        f8_fb = (torch.where(fr8 > 0.5, tr8, torch.full_like(tr8, math.inf)).min(),
                                 tr8.gather(0, top8)[0], fr8.gather(0, top8)[0])
        f = flags[r0:r1]
        compact = self.cfg.screen_compact and r1 > r0
        if compact:
            # Near the boundary a row tile holds proven and unproven rows side by side (C2: 18 k unproven rows spread over
            # 237 tiles = 61 k rows).  The unproven rows of the span are gathered into dense tiles -- index list, count and
            # tile count all stay on the device -- and go through the full-precision pass on their own; ``cap`` (from the
            # count this block had in the previous call) only sizes the buffers: rows beyond it stay on the tile-list path.
            span = r1 - r0
            hint_rows = plan_tiles[1] if isinstance(plan_tiles, tuple) else -1
            cap = min(span, max(2 * rpt, (2 * hint_rows + rpt) if hint_rows >= 0 else span // 4))
            cap = (cap + rpt - 1) // rpt * rpt
            open_span = f == 0
            pos = torch.cumsum(open_span, 0, dtype=torch.int64) - 1           # dense position of every unproven row
            take = open_span & (pos < cap)
            idx_c = torch.zeros(cap + 1, dtype=torch.int64, device=dev)       # slot ``cap`` collects everything not taken
            idx_c.scatter_(0, torch.where(take, pos, torch.full_like(pos, cap)), torch.arange(span, device=dev))
            idx_c = idx_c[:cap]
            n_c = (pos[-1] + 1).clamp(max=cap)
            n_tiles_c = ((n_c + rpt - 1) // rpt).to(torch.int32).reshape(1)
            # what the tile-list launch still has to cover: the unscreened rows and any unproven row beyond ``cap``
            flags_u = flags.clone()
            flags_u[r0:r1] = (~(open_span & ~take)).to(torch.uint8)
            tile_list, n_listed = be.screen_tile_list(flags_u, rpt)
        else:
            tile_list, n_listed = be.screen_tile_list(flags, rpt)
        with ph("screen: full pass over listed tiles"):
            plan_u = plan_tiles[0] if isinstance(plan_tiles, tuple) else plan_tiles
            parts = self._local_partials(prep, rows, inv_temp, aux, precision, row_tiles=tile_list, n_row_tiles=tiles,
                                         n_row_tiles_dev=n_listed, plan_row_tiles=plan_u)
            out, argmin = self._merge(parts, inv_temp)
            if compact:
                prep_c = {k: (v[r0:r1].index_select(0, idx_c) if isinstance(v, Tensor) else v) for k, v in prep.items()}
                inv_temp_c = inv_temp[r0:r1].index_select(0, idx_c)
                cap_tiles = cap // rpt
                iota = torch.arange(cap_tiles, dtype=torch.int32, device=dev)
                parts_c = self._local_partials(prep_c, cap, inv_temp_c, aux, precision, row_tiles=iota, n_row_tiles=cap_tiles,
                                               n_row_tiles_dev=n_tiles_c,
                                               plan_row_tiles=max(1, (hint_rows + rpt - 1) // rpt) if hint_rows >= 0 else 0)
                out_c, arg_c = self._merge(parts_c, inv_temp_c)
                back = pos.clamp(min=0, max=cap - 1)
                out[:, r0:r1] = torch.where(take, out_c.index_select(1, back), out[:, r0:r1])
                argmin[r0:r1] = torch.where(take, arg_c.index_select(0, back), argmin[r0:r1])
        with ph("screen: closed form"):
            y_hi, y_lo = ds.split()
            be.screen_finalize(flags, arg1, ds.d, (prep["hi"], prep["lo"], prep["inv_scale"]), prep["norms"],
                               (y_hi, None if precision == "f16x2" else y_lo), 1.0 / ds.scale, ds.y_norm, aux,
                               ds.index_offset, ds.n, ds.n_total, out, argmin)
            if self.world > 1:
                import torch.distributed as dist
                both = torch.stack([out[_cabi.OUT_E_MIN], -out[_cabi.OUT_AUX_MEAN]])
                dist.all_reduce(both, op=dist.ReduceOp.MIN, group=self.group)
                out[_cabi.OUT_E_MIN].copy_(both[0])
                out[_cabi.OUT_AUX_MEAN].copy_(-both[1])
        open_rows = f == 0
        frac, t_run = self._open_runs(open_rows, temp_rows[r0:r1], per_temp)
        t_open = torch.where(frac > 0.5, t_run, torch.full_like(t_run, math.inf)).min()
        # the highest screened temperature and how it fared.  (Indexing with the 0-dim result of argmax() would read it
        # back -- torch turns a 0-dim integer tensor index into .item() -- and stall the host once per block: gather.)
        top = t_run.argmax().reshape(1)
        nan = torch.full((), math.nan, dtype=torch.float64, device=dev)
        feedback = torch.stack([n_listed[0].to(torch.float64), t_open.to(torch.float64), open_rows.sum().to(torch.float64),
                                (nl8[0] if nl8 is not None else n_listed.new_full((1,), -1)[0]).to(torch.float64),
                                t_run.gather(0, top)[0].to(torch.float64), frac.gather(0, top)[0].to(torch.float64)]
                               + ([v.to(torch.float64) for v in f8_fb] if f8_fb is not None else [nan, nan, nan]))
        return out, argmin, feedback

    def stats_block(self, src: Tensor, rows: int, temp_rows: Tensor, *, noise: Optional[Tensor] = None,
                    sigma: Optional[Tensor] = None, post: Optional[Tensor] = None, aux: Optional[Tensor] = None,
                    prep: Optional[dict] = None, screen: bool = False, ascending: bool = True, wide: bool = True,
                    per_temp: int = 1):
        """Statistics for ``rows`` query rows; returns (out (8, rows), argmin (rows,)) device tensors.
        Query row r is  (noise[r]*sigma[r] + src[r % len(src)]) * post[r]  (noise/post optional); ``prep`` passes
        already prepared operands (rank-sliced preparation of a sharded run) instead."""
        precision = self.precision()
        ph = getattr(self.backend, "phase", None)
        if ph is None:
            import contextlib
            ph = lambda _n: contextlib.nullcontext()      # noqa: E731  (test doubles without phase timing)
        inv_temp = (1.0 / temp_rows.to(torch.float32)).contiguous()
        if prep is None:
            with ph("prepare"):
                prep = self._prepare(src, rows, noise, sigma, post, precision, want_x=False)
        if screen and self.screening_usable():
            with ph("fused"):
                return self._screened_block(prep, rows, temp_rows.to(torch.float32), inv_temp, aux, precision, ascending, wide,
                                            per_temp)
        self.screen_report["rows_unscreened"] += rows
        with ph("fused"):
            parts = self._local_partials(prep, rows, inv_temp, aux, precision)
        with ph("merge"):
            return self._merge(parts, inv_temp)

    @staticmethod
    def _empty_stats(shape: tuple, dev: torch.device) -> dict:
        res = {k: torch.empty(shape, dtype=torch.float32, device=dev) for k in STAT_KEYS}
        res["argmin"] = torch.empty(shape, dtype=torch.int64, device=dev)
        return res

    def stats(self, x: Tensor, temp_rows: Tensor, aux: Optional[Tensor] = None) -> dict:
        """Per-row Boltzmann statistics of explicit queries x (M, ...) at per-row temperatures."""
        dev = self.backend.device
        xf = _flat2d(x).to(device=dev, dtype=torch.float32).contiguous()
        temp_rows = temp_rows.to(device=dev, dtype=torch.float32).reshape(-1).expand(xf.shape[0]).contiguous()
        outs, idxs = [], []
        aux = self._local_aux(aux)
        if xf.shape[0] == 0:
            return self._empty_stats((0,), dev)
        step = self.rows_per_block()
        for r0 in range(0, xf.shape[0], step):
            r1 = min(xf.shape[0], r0 + step)
            o, i = self.stats_block(xf[r0:r1], r1 - r0, temp_rows[r0:r1], aux=aux)
            outs.append(o)
            idxs.append(i)
        out = torch.cat(outs, dim=1) if len(outs) > 1 else outs[0]
        res = {k: out[j] for j, k in enumerate(STAT_KEYS)}
        res["argmin"] = torch.cat(idxs) if len(idxs) > 1 else idxs[0]
        return res

    def noised_stats(self, x0: Tensor, temp: Tensor, aux: Optional[Tensor] = None, noise_fn=None) -> dict:
        """Statistics of xt = randn * sqrt(T_i) + x0 for every temperature of ``temp``.

        The noise is drawn with one ``torch.randn(*x0.shape, device=...)`` per temperature in schedule
        order -- the reference's RNG stream (utils/stats.py:74, :273).  With a ``query_group`` this rank draws and
        evaluates temperatures q_rank, q_rank + q_world, ... only (same stream: each draw sits at its own Philox offset)
        and the results are all-gathered.  Returns tensors of shape (n_T, B)."""
        dev = self.backend.device
        b = x0.shape[0]
        x0f = _flat2d(x0).to(device=dev, dtype=torch.float32).contiguous()
        temp = temp.to(device=dev, dtype=torch.float32).reshape(-1)
        n_t = temp.shape[0]
        aux = self._local_aux(aux)
        if b == 0 or n_t == 0:
            return self._empty_stats((n_t, b), dev)
        qw, qr = self.q_world, self.q_rank
        temp_mine = temp[qr::qw].contiguous()                 # temperature k of mine is number qr + k*qw of the schedule
        n_mine = temp_mine.shape[0]
        t_per_block = max(1, self.rows_per_block() // b)
        outs, idxs = [], []
        if noise_fn is None and PosteriorEngine.noise_hook is not None:
            noise_fn = lambda i: PosteriorEngine.noise_hook(i, tuple(x0.shape), dev)     # noqa: E731
        draw = noise_fn
        fused = draw is None and self._fused_noise_usable(x0, dev)
        if (self.world > 1 or qw > 1) and draw is None:
            self._sync_generator(dev)            # every rank continues rank 0's stream: 16 bytes instead of the noise
        x0_absmax = self.backend.row_absmax(x0f) if fused and self.precision() != "exact" else None
        # Where the draws sit in the generator's stream.  CUDA: draw number i of the call starts at Philox offset
        # base + i*step (step measured once per shape), so a rank can jump straight to its own.  CPU generator (test
        # double of the backend): no random access -- the draws of other ranks are made and dropped.
        on_cuda = dev.type == "cuda"
        gen = base = step_off = None
        if draw is None and on_cuda:
            gen = self._cuda_generator(dev)
            step_off = self._randn_offset_step(tuple(x0.shape), dev)
            base = gen.get_offset()
        drawn = 0                                 # CPU stream position, in draws of this call

        def draw_into(dst: Tensor, i: int) -> None:
            nonlocal drawn
            if draw is not None:
                dst.copy_(draw(i).reshape(b, -1))
            elif on_cuda:
                gen.set_offset(base + i * step_off)
                torch.randn(*x0.shape, device=dev, out=dst.view(x0.shape))    # the generator calls of torch.randn(*shape)
            else:
                while drawn < i:
                    torch.randn(*x0.shape)
                    drawn += 1
                torch.randn(*x0.shape, out=dst.view(x0.shape))
                drawn += 1

        # Screening policy: a block is screened while its lowest temperature is below the lowest temperature at which a
        # row has failed the certificate so far in this call (schedules run from low to high noise; a block above that
        # mark would pay the one-product pass for nothing).  After a screened block that certifies nothing the
        # next attempt waits for a block that reaches a quarter of its lowest temperature (schedules that start at the
        # high-noise end pay a logarithmic number of failed attempts).
        screen_on = self.screening_usable()
        temp_host = temp_mine.detach().cpu() if screen_on else None
        self._screen_t_fail = math.inf
        self._screen_t_retry = math.inf
        self._screen_f8_live = True
        # From the second call on the certifiable range is known (one dataset, one boundary: the lowest temperature at
        # which most rows stayed unproven, remembered across calls): a block screens its rows below that mark in one go
        # and leaves the rest, without probing and without reading anything back until the end of the call.
        prior = self._screen_prior if (screen_on and hasattr(self.backend, "screen_merge_stage")) else None
        rpt = getattr(self.backend, "row_tile", None) or 128 * (self.cfg.cta_group or 2)
        pending = []                              # (feedback scalars on the device, rows screened, tiles screened, block key)
        ph = getattr(self.backend, "phase", None)
        if ph is None:
            import contextlib
            ph = lambda _n: contextlib.nullcontext()  # noqa: E731
        for k0 in range(0, n_mine, t_per_block):
            k1 = min(n_mine, k0 + t_per_block)
            nb = k1 - k0
            tb = temp_mine[k0:k1]
            screen = screen_on and float(temp_host[k0:k1].min()) < min(self._screen_t_fail, self._screen_t_retry)
            asc = not screen or bool(temp_host[k0] <= temp_host[k1 - 1])
            wide = screen and float(temp_host[k0:k1].max()) > 8.0 * float(temp_host[k0:k1].min())
            t_rows = tb.repeat_interleave(b)
            th = temp_host[k0:k1] if screen_on else None
            monotone = screen_on and (bool((th[1:] >= th[:-1]).all()) or bool((th[1:] <= th[:-1]).all()))
            if prior is not None and monotone:
                # temperatures of the block at or below the cut sit at one end of it
                below = int((th < prior).sum())
                span = None
                if below > 0:
                    if bool(th[0] <= th[-1]):
                        span = (0, min(nb * b, -(-below * b // rpt) * rpt))
                    else:
                        span = ((nb - below) * b // rpt * rpt, nb * b)
                if span is not None and span[1] - span[0] >= min(rpt, nb * b):
                    with ph("noise+prepare" if fused else "noise"):
                        if fused:
                            prep = self._fused_prepare(gen.initial_seed(), base + (qr + k0 * qw) * step_off, qw * step_off, x0f,
                                                       tb, x0_absmax)
                        else:
                            noise = torch.empty(nb, b, self.ds.d, dtype=torch.float32, device=dev)
                            for i in range(nb):
                                draw_into(noise[i], qr + (k0 + i) * qw)
                            prep = self._prepare(x0f, nb * b, noise.view(nb * b, -1), t_rows.sqrt(), None, self.precision(), False)
                    key = (n_t, b, k0)
                    # the E4M3 first stage has a mark of its own (lower: 4 significant bits): rows of the span above it start
                    # at the fp16 stage instead of paying for an E4M3 pass that is known to leave them unproven
                    f8_span = None
                    if self._screen_prior_f8 is not None:
                        below8 = int((th < self._screen_prior_f8).sum())
                        if below8 == 0:
                            f8_span = (span[0], span[0])
                        elif bool(th[0] <= th[-1]):
                            f8_span = (span[0], min(span[1], -(-below8 * b // rpt) * rpt))
                        else:
                            f8_span = (max(span[0], (nb - below8) * b // rpt * rpt), span[1])
                    o, i, fb = self._screened_block_prior(prep, nb * b, t_rows, (1.0 / t_rows).contiguous(), aux,
                                                          self.precision(), span, self._screen_hint.get(key, 0), per_temp=b,
                                                          f8_rows=f8_span)
                    s8 = f8_span if f8_span is not None else span
                    pending.append((fb, span[1] - span[0], (span[1] - span[0] + rpt - 1) // rpt, key, nb * b,
                                    (s8[1] - s8[0] + rpt - 1) // rpt))
                    outs.append(o)
                    idxs.append(i)
                    continue
                screen = False                    # nothing of this block lies in the certifiable range
            if fused:
                with ph("noise+prepare"):
                    prep = self._fused_prepare(gen.initial_seed(), base + (qr + k0 * qw) * step_off, qw * step_off, x0f, tb,
                                               x0_absmax)
                o, i = self.stats_block(x0f, nb * b, t_rows, aux=aux, prep=prep, screen=screen, ascending=asc, wide=wide,
                                        per_temp=b)
            else:
                with ph("noise"):
                    noise = torch.empty(nb, b, self.ds.d, dtype=torch.float32, device=dev)
                    for i in range(nb):
                        draw_into(noise[i], qr + (k0 + i) * qw)
                o, i = self.stats_block(x0f, nb * b, t_rows, noise=noise.view(nb * b, -1), sigma=t_rows.sqrt(), aux=aux,
                                        screen=screen, ascending=asc, wide=wide, per_temp=b)
            outs.append(o)
            idxs.append(i)
        if screen_on:
            self._screen_feedback(pending, float(temp_host.max()) if n_mine > 0 else math.inf, rpt)
        if draw is None:                          # leave the generator where a single-GPU run would
            if on_cuda:
                gen.set_offset(base + n_t * step_off)
            else:
                while drawn < n_t:
                    torch.randn(*x0.shape)
                    drawn += 1
        if n_mine > 0:
            out = (torch.cat(outs, dim=1) if len(outs) > 1 else outs[0]).view(len(STAT_KEYS), n_mine, b)
            arg = (torch.cat(idxs) if len(idxs) > 1 else idxs[0]).view(n_mine, b)
        else:
            out = torch.empty(len(STAT_KEYS), 0, b, dtype=torch.float32, device=dev)
            arg = torch.empty(0, b, dtype=torch.int64, device=dev)
        if qw > 1:
            out, arg = self._gather_temperatures(out, arg, n_t)
        res = {k: out[j] for j, k in enumerate(STAT_KEYS)}
        res["argmin"] = arg
        return res

    def _screen_feedback(self, pending: list, t_max: float, rpt: int) -> None:
        """End of a noised_stats call: read the screening counts of its blocks (one copy for all of them) and move the
        remembered boundary: the lowest temperature at which a row stayed unproven; if every screened row was proven the
        mark moves up by the factor the next call will try."""
        rep = self.screen_report
        if pending:
            vals = torch.stack([p[0] for p in pending]).cpu()
            t_fail, t_top, top_open = math.inf, 0.0, 1.0
            t_fail8, t_top8, top_open8 = math.inf, 0.0, 1.0
            for (_, rows_s, tiles_s, key, rows_blk, tiles8), v in zip(pending, vals.tolist()):
                n_l, t_open, n_open, n8 = int(v[0]), float(v[1]), int(v[2]), int(v[3])
                if not math.isnan(v[6]):                      # this block ran the E4M3 stage: how its own verdicts fell
                    t_fail8 = min(t_fail8, float(v[6]))
                    if v[7] > t_top8:
                        t_top8, top_open8 = float(v[7]), float(v[8])
                unscreened_tiles = (rows_blk - rows_s + rpt - 1) // rpt
                rep["rows_screened"] += rows_s
                rep["rows_certified"] += rows_s - n_open
                rep["rows_unscreened"] += rows_blk - rows_s
                rep["tiles_screened"] += tiles_s
                rep["tiles_full_pass"] += max(0, n_l - unscreened_tiles) + ((n_open + rpt - 1) // rpt if self.cfg.screen_compact else 0)
                if n8 >= 0:
                    rep["f8_tiles_screened"] = rep.get("f8_tiles_screened", 0) + tiles8
                    rep["f8_tiles_left"] = rep.get("f8_tiles_left", 0) + n8
                self._screen_hint[key] = (n_l + 2, n_open)    # schedule / buffer hints for the same block of the next call
                t_fail = min(t_fail, t_open)
                if v[4] > t_top:
                    t_top, top_open = float(v[4]), float(v[5])
            # a temperature failed by majority: the mark moves there (strictly below it is screened next time); nothing
            # failed and the highest screened temperature was proven almost entirely: try 1.25x further next time
            if math.isfinite(t_fail):
                self._screen_prior = t_fail
            elif top_open < 0.2:
                self._screen_prior = max(self._screen_prior, min(1.25 * t_top, 2.0 * t_max))
            # the E4M3 stage's mark moves by the same rule on the stage's own verdicts
            if math.isfinite(t_fail8):
                self._screen_prior_f8 = t_fail8
            elif t_top8 > 0.0 and top_open8 < 0.2:
                self._screen_prior_f8 = max(self._screen_prior_f8 or 0.0, min(1.25 * t_top8, 2.0 * t_max))
        elif self._screen_prior is None and math.isfinite(self._screen_t_fail):
            self._screen_prior = self._screen_t_fail          # first call: the probing found the boundary
        elif self._screen_prior is None and self.screen_report["rows_screened"] > 0 and not math.isfinite(self._screen_t_fail):
            self._screen_prior = 2.0 * t_max                  # everything proven: screen the whole schedule next time

    def _gather_temperatures(self, out: Tensor, arg: Tensor, n_t: int):
        """(8, n_mine, B) statistics and (n_mine, B) arg-mins of this rank's temperatures -> those of the whole schedule,
        on every rank of the query group (one all-gather of each; ranks are padded to the same number of temperatures)."""
        import torch.distributed as dist
        qw = self.q_world
        n_max = (n_t + qw - 1) // qw
        b = out.shape[2]
        k = out.shape[0]
        send = torch.zeros(k + 2, n_max, b, dtype=torch.float32, device=out.device)
        send[:k, :out.shape[1]] = out
        send[k:, :arg.shape[0]] = arg.contiguous().view(torch.int32).view(arg.shape[0], b, 2).permute(2, 0, 1).view(torch.float32)
        recv = torch.empty((qw,) + tuple(send.shape), dtype=torch.float32, device=out.device)
        dist.all_gather_into_tensor(recv.view(qw * (k + 2), n_max, b), send, group=self.query_group)
        full = torch.empty(k, n_t, b, dtype=torch.float32, device=out.device)
        full_arg = torch.empty(n_t, b, dtype=torch.int64, device=out.device)
        for j in range(qw):
            cnt = len(range(j, n_t, qw))
            full[:, j::qw] = recv[j, :k, :cnt]
            halves = recv[j, k:, :cnt].view(torch.int32).permute(1, 2, 0).contiguous()       # (cnt, b, 2)
            full_arg[j::qw] = halves.view(torch.int64).view(cnt, b)
        return full, full_arg

    # -- the reference's RNG stream ---------------------------------------------------------------------
    _RANDN_OFFSETS: dict = {}

    @staticmethod
    def _randn_offset_step(shape: tuple, dev: torch.device) -> int:
        """By how much one ``torch.randn(*shape, device=dev)`` advances the Philox offset of the CUDA generator
        (a function of numel and the device; measured on a scratch generator, never guessed)."""
        key = (int(math.prod(shape)), str(dev))
        step = PosteriorEngine._RANDN_OFFSETS.get(key)
        if step is None:
            g = torch.Generator(device=dev)
            g.manual_seed(0)
            o0 = g.get_offset()
            torch.randn(*shape, device=dev, generator=g)
            step = g.get_offset() - o0
            PosteriorEngine._RANDN_OFFSETS[key] = step
        return step

    def _sync_generator(self, dev: torch.device) -> None:
        """Give every rank of the grid rank 0's generator state (seed and offset), so that each can draw its part of the
        one noise stream a single-GPU run would draw: first along the dataset group, then along the query group."""
        if not self.cfg.sync_noise:
            return
        import torch.distributed as dist
        cuda = dev.type == "cuda"
        state = (torch.cuda.get_rng_state(dev) if cuda else torch.get_rng_state()).to(dev)
        for grp in (self.group, self.query_group):
            if grp is not None and dist.get_world_size(grp) > 1:
                dist.broadcast(state, src=dist.get_global_rank(grp, 0), group=grp)
        if cuda:
            torch.cuda.set_rng_state(state.cpu(), dev)
        else:
            torch.set_rng_state(state.cpu())

    @staticmethod
    def _cuda_generator(dev: torch.device):
        torch.cuda.init()                       # torch.cuda.default_generators is filled by the lazy initialisation
        return torch.cuda.default_generators[dev.index if dev.index is not None else torch.cuda.current_device()]

    _FUSED_NOISE_OK: dict = {}

    def _fused_noise_usable(self, x0: Tensor, dev: torch.device) -> bool:
        """The in-kernel Philox path reproduces torch.randn bit for bit on this device (checked once, against
        torch.randn itself, for two shapes, a non-zero offset and consecutive draws) and the shape is one it covers."""
        if not self.cfg.fused_noise or dev.type != "cuda" or not hasattr(self.backend, "noised_rows_philox"):
            return False
        if x0.numel() >= 2 ** 31 or (self.precision() != "exact" and self.ds.d % 8 != 0):
            return False
        key = str(dev)
        ok = PosteriorEngine._FUSED_NOISE_OK.get(key)
        if ok is None:
            ok = True
            try:
                for shape in ((48, 3072), (5, 331)):
                    g = torch.Generator(device=dev)
                    g.manual_seed(20240229)
                    g.set_offset(24)
                    step = self._randn_offset_step(shape, dev)
                    want = torch.stack([torch.randn(*shape, device=dev, generator=g) for _ in range(3)])
                    got = self.backend.noised_rows_philox(20240229, 24, step, torch.zeros(shape, device=dev),
                                                          torch.ones(3, device=dev), want_x=True, want_split=False)["x"]
                    ok = ok and torch.equal(got.view(3, *shape), want) and g.get_offset() == 24 + 3 * step
            except Exception:       # noqa: BLE001  (any surprise in torch's generator API: use torch.randn itself)
                ok = False
            PosteriorEngine._FUSED_NOISE_OK[key] = ok
        return ok

    def _fused_prepare(self, seed: int, offset: int, step: int, x0f: Tensor, temps: Tensor, x0_absmax) -> dict:
        tensor = self.precision() != "exact"
        return self.backend.noised_rows_philox(seed, offset, step, x0f, temps.sqrt().contiguous(), x0_absmax=x0_absmax,
                                               want_x=not tensor, want_split=tensor)

    # -- posterior mean ---------------------------------------------------------------------------
    def posterior_mean(self, x: Tensor, temp_rows: Tensor, post: Optional[Tensor] = None,
                       values: Optional[Tensor] = None, temp_bounds: Optional[tuple] = None, scatter: bool = False,
                       frozen: Optional[tuple] = None) -> Tensor:
        """x0_hat[r] = sum_j p_rj y_j with p ~ exp(-||x_r*post_r - y_j||^2 / (2 T_r)).  Returns (M, d).
        ``values`` (N, dv) replaces y_j in the weighted sum (posterior mean of arbitrary per-point vectors).

        Nothing in here waits for the device: which rows are delta posteriors, which row tiles a screening stage left
        unproven and which tiles have to be contracted are tile lists whose lengths stay in device memory
        (pdm_stats_args.n_row_tiles_dev and the *_tiles entry points), so a sampling loop enqueues step after step.
        ``temp_bounds`` = (lowest, highest) temperature of the call as host floats (the sampler knows them); without it
        an engine with screening on reads the two numbers back once per call.
        ``scatter`` (row-sharded dataset): the shards' partial sums are combined with a reduce-scatter instead of an
        all-reduce and the call returns only this rank's slice of the rows, ceil(M / world) of them starting at
        rank * ceil(M / world) (zero rows beyond M) -- a sampler whose ranks each own a slice of the trajectories moves
        M d / world floats per rank instead of M d.
        ``frozen`` = (screen?, E4M3 stage?, pinned int32[2] buffer): the screening decisions are the caller's and the counts
        are copied to the caller's buffer -- the form a CUDA-graph capture of a sampling step needs (IdealSampler)."""
        dev = self.backend.device
        be = self.backend
        ds = self.ds
        xf = _flat2d(x).to(device=dev, dtype=torch.float32).contiguous()
        m = xf.shape[0]
        temp_rows = temp_rows.to(device=dev, dtype=torch.float32).reshape(-1).expand(m).contiguous()
        if post is not None:
            post = post.to(device=dev, dtype=torch.float32).reshape(-1).expand(m).contiguous()
        precision = self.precision()
        tensor = precision != "exact"
        vt = None
        if values is not None:
            values = _flat2d(values).to(device=dev, dtype=torch.float32).contiguous()
            if values.shape[0] != ds.n:
                raise PdmError("values must have one row per dataset row")
            if tensor:
                vscale = pow2_scale_for(float(be.absmax(values).item()))
                vt = be.transpose_split(values, vscale) + (vscale,)
        src = ds.y if values is None else values
        per = (m + self.world - 1) // self.world
        m_alloc = per * self.world if (scatter and self.world > 1) else m
        out = torch.empty(m_alloc, src.shape[1], dtype=torch.float32, device=dev)
        if m_alloc > m:
            out[m:].zero_()
        step = max(128, min(m, self.cfg.max_energy_bytes // (ds.n * 8)))
        rpt = getattr(be, "row_tile", None) or 128 * (self.cfg.cta_group or 2)          # the fused kernel's row tile
        tile_ops = tensor and hasattr(be, "delta_tile_list")                             # device-side tile lists available
        ph = getattr(be, "phase", None)
        if ph is None:
            import contextlib
            ph = lambda _n: contextlib.nullcontext()      # noqa: E731
        screen_on = self.screening_usable() and tile_ops and m > 0
        if screen_on and frozen is None:
            self._pm_poll()
            if temp_bounds is None:
                temp_bounds = tuple(float(v) for v in torch.stack([temp_rows.min(), temp_rows.max()]).cpu())
        for r0 in range(0, m, step):
            r1 = min(m, r0 + step)
            rows = r1 - r0
            tiles = (rows + rpt - 1) // rpt
            inv_temp = (1.0 / temp_rows[r0:r1]).contiguous()
            out_blk = out[r0:r1]
            with ph("prepare"):
                prep = self._prepare(xf[r0:r1], rows, None, None, None if post is None else post[r0:r1], precision, False)
            # Certified delta posteriors (the low-noise steps of a sampling trajectory): the one-product passes prove rows to
            # be deltas; only the row tiles with an unproven row get full-precision distances at all.  The marks move on the
            # counts of EARLIER calls (a failed attempt at temperature T is repeated below T/2, a success moves the mark up
            # to 1.5 T; they persist across calls: one dataset, one boundary).
            listed = None                    # (tile list, device-side length) of the full-precision pass; None = every tile
            flags_s = arg_s = None
            attempt = screen_on and (frozen[0] if frozen is not None else temp_bounds[1] < self._pm_screen_t)
            if attempt:
                with ph("screen"):
                    use_f8 = frozen[1] if frozen is not None else temp_bounds[1] < self._pm_f8_t
                    flags_s, arg_s, tl, nl, _, nl8 = self._screen_cascade_async(prep, rows, inv_temp, use_f8)
                    listed = (tl, nl)
                    if frozen is None:
                        self._pm_report(nl, nl8, tiles, rows, temp_bounds[0], temp_bounds[1])
                    else:
                        frozen[2].copy_(torch.cat([nl.reshape(1), nl8.reshape(1) if nl8 is not None else nl.new_full((1,), -1)]),
                                        non_blocking=True)
            with ph("prepare"):
                energy = torch.empty(rows, ds.n, dtype=torch.float32, device=dev)
            with ph("fused+energy"):
                kw = {} if listed is None else dict(row_tiles=listed[0], n_row_tiles=tiles, n_row_tiles_dev=listed[1])
                parts = self._local_partials(prep, rows, inv_temp, None, precision, energy_out=energy, energy_mult=1.0, **kw)
            with ph("merge"):
                st, amin = self._merge(parts, inv_temp)
            if flags_s is not None:
                y_hi, y_lo = ds.split()                   # proven rows: l = 1, arg-min of the screening passes
                be.screen_finalize(flags_s, arg_s, ds.d, (prep["hi"], prep["lo"], prep["inv_scale"]), prep["norms"],
                                   (y_hi, None if precision == "f16x2" else y_lo), 1.0 / ds.scale, ds.y_norm, None,
                                   ds.index_offset, ds.n, ds.n_total, st, amin)
            e_min, l = st[_cabi.OUT_E_MIN], st[_cabi.OUT_L]
            # Rows whose posterior is a delta to fp32 resolution (every other weight sums to <= 2^-23): the mean IS the
            # nearest training point -- a gather (on the shard that owns it) instead of weights + second contraction.
            dflags = todo = None
            if (self.cfg.delta_shortcut or flags_s is not None) and hasattr(be, "delta_tile_list"):
                with ph("delta rows"):
                    dflags, tl2, nl2 = be.delta_tile_list(l, rpt)
                    todo = (tl2, nl2)
            if tensor:
                with ph("weights"):
                    tkw = {} if todo is None else dict(tiles=(todo[0], rpt, tiles, todo[1]))
                    p_hi, p_lo = be.weights_from_energy(energy, e_min, l, inv_temp, split=True, **tkw)
                with ph("gemm2"):
                    yt_hi, yt_lo, yscale = vt if vt is not None else (ds.transposed_split() + (ds.scale,))
                    if vt is None and precision == "f16x2":
                        yt_lo = None                      # lattice dataset: weights_hi.Y + weights_lo.Y is all there is
                    tkw = {} if todo is None else dict(tiles=(todo[0], tiles, todo[1]))
                    be.split_gemm(p_hi, p_lo, yt_hi, yt_lo, ds.n, (1.0 / 16384.0) / yscale, out=out_blk,
                                  cta_group=self.cfg.cta_group, **tkw)
            else:
                with ph("weights"):
                    p = be.weights_from_energy(energy, e_min, l, inv_temp, split=False)
                with ph("gemm2"):
                    be.weighted_mean_exact(p, src, out=out_blk)
            if dflags is not None:
                with ph("delta rows"):
                    be.gather_rows(src, amin, ds.index_offset, dflags, out_blk)
        if self.world > 1:
            import torch.distributed as dist
            if scatter:
                mine = torch.empty(per, out.shape[1], dtype=torch.float32, device=dev)
                dist.reduce_scatter_tensor(mine, out, group=self.group)
                return mine
            dist.all_reduce(out, group=self.group)
        return out

    def posterior_mean_backward(self, x: Tensor, temp_rows: Tensor, post: Optional[Tensor], grad_out: Tensor):
        """Vector-Jacobian product of ``posterior_mean`` (no ``values``): for upstream gradient g (M, d) returns
        (g_q (M, d), g_T (M,)), the gradients with respect to the VE query rows q = x * post and to the per-row
        temperature.  The energies are recomputed (nothing of size M x N is kept between forward and backward):
            s_j = y_j . g,  a = sum_j p_j s_j,   g_q = sum_j p_j (s_j - a) y_j / T = Cov_p(s, y) / T,
            g_T = sum_j p_j (s_j - a) e_j / T = Cov_p(s, e) / T        with e_j = (E_j - E_min) / T."""
        dev = self.backend.device
        ds = self.ds
        xf = _flat2d(x).to(device=dev, dtype=torch.float32).contiguous()
        g = _flat2d(grad_out).to(device=dev, dtype=torch.float32).contiguous()
        m = xf.shape[0]
        temp_rows = temp_rows.to(device=dev, dtype=torch.float32).reshape(-1).expand(m).contiguous()
        if post is not None:
            post = post.to(device=dev, dtype=torch.float32).reshape(-1).expand(m).contiguous()
        precision = self.precision()
        tensor = precision != "exact"
        g_q = torch.empty(m, ds.d, dtype=torch.float32, device=dev)
        g_t = torch.empty(m, dtype=torch.float32, device=dev)
        step = max(128, min(m, self.cfg.max_energy_bytes // (ds.n * 12)))      # energy, S and w tiles
        for r0 in range(0, m, step):
            r1 = min(m, r0 + step)
            rows = r1 - r0
            inv_temp = (1.0 / temp_rows[r0:r1]).contiguous()
            prep = self._prepare(xf[r0:r1], rows, None, None, None if post is None else post[r0:r1], precision, False)
            energy = torch.empty(rows, ds.n, dtype=torch.float32, device=dev)
            parts = self._local_partials(prep, rows, inv_temp, None, precision, energy_out=energy, energy_mult=1.0)
            st, _ = self._merge(parts, inv_temp)
            e_min, l = st[_cabi.OUT_E_MIN], st[_cabi.OUT_L]
            if tensor:
                y_hi, y_lo = ds.split()
                gp = self.backend.prepare_rows(g[r0:r1], rows, want_norms=False)
                sdot = self.backend.split_gemm(gp["hi"], gp["lo"], y_hi, None if precision == "f16x2" else y_lo, ds.d,
                                               1.0 / ds.scale, cta_group=self.cfg.cta_group)      # (rows, n) * 2^k_row
                w, sums = self._centred_weights(energy, sdot, e_min, l, inv_temp, gp["inv_scale"])
                wp = self.backend.prepare_rows(w, rows, want_norms=False)
                yt_hi, yt_lo = ds.transposed_split()
                acc = self.backend.split_gemm(wp["hi"], wp["lo"], yt_hi, None if precision == "f16x2" else yt_lo, ds.n,
                                              1.0 / ds.scale, cta_group=self.cfg.cta_group)
                acc = acc * wp["inv_scale"][:, None]
            else:
                sdot = self.backend.weighted_mean_exact(g[r0:r1], ds.transposed())
                w, sums = self._centred_weights(energy, sdot, e_min, l, inv_temp, None)
                acc = self.backend.weighted_mean_exact(w, ds.y)
            g_q[r0:r1] = acc * inv_temp[:, None]
            g_t[r0:r1] = sums[:, 1] * inv_temp
        if self.world > 1:                      # Cov_p(s, y) and Cov_p(s, e) are sums over the shards' dataset rows
            import torch.distributed as dist
            dist.all_reduce(g_q, group=self.group)
            dist.all_reduce(g_t, group=self.group)
        return g_q, g_t

    def _centred_weights(self, energy, sdot, e_min, l, inv_temp, s_scale):
        """w = p (s - a) and (a, sum w e).  With a row-sharded dataset a = sum_j p_j s_j runs over every shard: local
        sums first, all-reduce, then the centring with the global a."""
        if self.world == 1:
            return self.backend.denoiser_backward_weights(energy, sdot, e_min, l, inv_temp, s_scale)
        import torch.distributed as dist
        _, sums = self.backend.denoiser_backward_weights(energy, sdot, e_min, l, inv_temp, s_scale)
        a = sums[:, 0].contiguous()
        dist.all_reduce(a, group=self.group)
        return self.backend.denoiser_backward_weights(energy, sdot, e_min, l, inv_temp, s_scale, a_in=a)

    # -- nearest neighbours -----------------------------------------------------------------------------
    TOPK_MAX = _cabi.TOPK_SLOTS

    def nearest(self, x: Tensor, k: int, refine: bool = True):
        """The k smallest squared distances ||x_b - y_j||^2 of every query to this dataset shard and the global indices of
        those points: (values (M, k) ascending, indices (M, k) int64, ties by lower index; +inf / -1 beyond the shard's
        size).  Tensor path, k <= 8: the top-k epilogue of the fused kernel -- nothing of size M x N reaches HBM.  Otherwise
        (d < 64, k > 8): dense distance tiles + the selection kernel.  ``refine``: the selection runs on the norm expansion
        of utils/distance.py:21 (fp32 round-off 2^-24 (|x|^2 + |y|^2)); the candidates -- all 8 slots, so that near-ties
        misordered by that round-off are still among them -- are then re-evaluated directly in fp64 and re-sorted, which is
        the accuracy of the float64 search sklearn runs for utils/stats.py:50-60, 138-146."""
        dev = self.backend.device
        be = self.backend
        ds = self.ds
        xf = _flat2d(x).to(device=dev, dtype=torch.float32).contiguous()
        m = xf.shape[0]
        precision = self.precision()
        can_refine = refine and hasattr(be, "refine_neighbours")
        if precision != "exact" and k <= self.TOPK_MAX and hasattr(be, "topk_merge"):
            ks = self.TOPK_MAX if can_refine else k                     # candidates kept for the refinement
            vals = torch.empty(m, ks, dtype=torch.float32, device=dev)
            idx = torch.empty(m, ks, dtype=torch.int64, device=dev)
            step = self.rows_per_block()
            y_hi, y_lo = ds.split()
            for r0 in range(0, m, step):
                r1 = min(m, r0 + step)
                prep = self._prepare(xf[r0:r1], r1 - r0, None, None, None, precision, False)
                tv, ti = be.posterior_stats(precision=precision, M=r1 - r0, N=ds.n, d=ds.d, q_norm=prep["norms"], y_norm=ds.y_norm,
                                            inv_temp=None, q_split=(prep["hi"], prep["lo"], prep["inv_scale"]),
                                            y_split=(y_hi, None if precision == "f16x2" else y_lo), y_inv_scale=1.0 / ds.scale,
                                            index_offset=ds.index_offset, n_splits=self.cfg.n_splits, m_group=self.cfg.m_group,
                                            cta_group=self.cfg.cta_group, topk=True)
                v, i = be.topk_merge(tv, ti, ds.index_offset, ks)
                vals[r0:r1], idx[r0:r1] = v, i
        else:
            ks = min(max(k, self.TOPK_MAX) if (can_refine and k <= self.TOPK_MAX) else k, ds.n)
            vals = torch.full((m, max(k, ks)), float("inf"), dtype=torch.float32, device=dev)
            idx = torch.full((m, max(k, ks)), -1, dtype=torch.int64, device=dev)
            step = max(1, (1 << 30) // (4 * max(1, ds.n)))
            for r0 in range(0, m, step):
                d2 = self.pairwise_sqdist(xf[r0:r0 + step])
                if hasattr(be, "topk_smallest"):
                    v, i = be.topk_smallest(d2, ks)
                else:                                                     # CPU test double
                    v, i = torch.topk(d2, ks, dim=1, largest=False)
                vals[r0:r0 + step, :ks] = v
                idx[r0:r0 + step, :ks] = i + ds.index_offset
        if can_refine and vals.shape[1] <= self.TOPK_MAX and m > 0:
            be.refine_neighbours(xf, ds.y, ds.index_offset, vals, idx)
        return vals[:, :k].contiguous(), idx[:, :k].contiguous()

    # -- dense distances (callers of compute_pw_dist_sqr index / min / scatter the matrix) -----------
    def pairwise_sqdist(self, x: Tensor) -> Tensor:
        """Dense (M, N) squared distances ||x_b - y_j||^2 in the reference's op order (utils/distance.py:21)."""
        dev = self.backend.device
        xf = _flat2d(x).to(device=dev, dtype=torch.float32).contiguous()
        m = xf.shape[0]
        precision = self.precision()
        out = torch.empty(m, self.ds.n, dtype=torch.float32, device=dev)
        prep = self._prepare(xf, m, None, None, None, precision, False)
        self._local_partials(prep, m, None, None, precision, energy_out=out, energy_mult=2.0, want_partials=False)
        return out
