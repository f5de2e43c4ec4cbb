"""Tensor-level wrapper over the C ABI: torch owns the device memory and the stream, the library does
the arithmetic.  One method per entry point of include/pdm_b200.h; no computation happens in torch here.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
from torch import Tensor

from . import _cabi
from ._cabi import PdmError, StatsArgs, check

PRECISIONS = {"exact": _cabi.PREC_EXACT_F32, "f16x3": _cabi.PREC_F16X3, "f16x1": _cabi.PREC_F16X1,
              "f16x2": _cabi.PREC_F16X2, "f8x1": _cabi.PREC_F8X1}


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _ld(t: Tensor) -> int:
    """Leading dimension in elements (a single-row matrix may carry any stride(0))."""
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[-1])


def _round_up(a: int, b: int) -> int:
    return (a + b - 1) // b * b


class CudaBackend:
    """Dispatches into libpdm_b200.so.  Raises PdmError if the library or a CUDA device is missing."""

    name = "cuda"

    def __init__(self, device: Optional[torch.device] = None):
        self.lib = _cabi.load()
        if not torch.cuda.is_available():
            raise PdmError("pdm_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        sm, major, minor = C.c_int(), C.c_int(), C.c_int()
        check(self.lib.pdm_device_info(self.device.index or 0, C.byref(sm), C.byref(major), C.byref(minor)),
              "pdm_device_info")
        self.sm_count, self.cc = sm.value, (major.value, minor.value)
        self.launches = 0          # kernels of ours launched through this backend (bench reports it)
        self.kernel_events = None  # bench hook: list of (start, end, pairs) CUDA events around the fused kernel

    # ---- optional phase timing (dev / bench): CUDA events on the launching stream ----------------
    def phase(self, name: str):
        """Context manager: when ``self.phase_events`` is a dict, records (start, end) events per phase."""
        backend = self

        class _Phase:
            def __enter__(self_inner):
                self_inner.on = getattr(backend, "phase_events", None) is not None
                if self_inner.on:
                    self_inner.e0 = torch.cuda.Event(enable_timing=True)
                    self_inner.e0.record()

            def __exit__(self_inner, *exc):
                if self_inner.on:
                    e1 = torch.cuda.Event(enable_timing=True)
                    e1.record()
                    backend.phase_events.setdefault(name, []).append((self_inner.e0, e1))
                return False

        return _Phase()

    def phase_totals(self) -> dict:
        ev = getattr(self, "phase_events", None) or {}
        return {k: sum(a.elapsed_time(b) for a, b in v) for k, v in ev.items()}

    # ---- helpers -------------------------------------------------------------------------------
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _f32(self, t: Tensor) -> Tensor:
        if t.device != self.device or t.dtype != torch.float32 or t.stride(-1) != 1:
            raise PdmError(f"expected a contiguous-row float32 tensor on {self.device}, got {t.dtype} on {t.device}")
        return t

    def supports_tensor_path(self) -> bool:
        return self.cc[0] == 10

    # ---- K1 ------------------------------------------------------------------------------------
    def row_norms(self, x: Tensor) -> Tensor:
        x = self._f32(x)
        out = torch.empty(x.shape[0], dtype=torch.float32, device=self.device)
        check(self.lib.pdm_row_norms_f32(x.data_ptr(), x.shape[0], x.shape[1], _ld(x), out.data_ptr(),
                                         self._stream()), "pdm_row_norms_f32")
        self.launches += 1
        return out

    def absmax(self, x: Tensor) -> Tensor:
        x = self._f32(x)
        out = torch.empty(1, dtype=torch.float32, device=self.device)
        check(self.lib.pdm_absmax_f32(x.data_ptr(), x.shape[0], x.shape[1], _ld(x), out.data_ptr(),
                                      self._stream()), "pdm_absmax_f32")
        self.launches += 1
        return out

    def lattice_residual(self, y: Tensor, scale: float) -> Tensor:
        """(max_j ||y_j - rint(y_j s)/s||^2 / ||y_j||^2,  max |rint(y s)|) as a 2-element device tensor."""
        y = self._f32(y)
        out = torch.empty(2, dtype=torch.float32, device=self.device)
        check(self.lib.pdm_lattice_residual_f32(y.data_ptr(), y.shape[0], y.shape[1], _ld(y), float(scale),
                                                out.data_ptr(), self._stream()), "pdm_lattice_residual_f32")
        self.launches += 1
        return out

    # ---- noised rows from torch's Philox stream ---------------------------------------------------
    def randn_launch_threads(self, numel: int) -> int:
        """Threads of the kernel torch.randn(numel) launches on this device (calc_execution_policy of
        ATen/native/cuda/DistributionTemplates.h): 256 * min(SMs * (maxThreadsPerSM // 256), ceil(numel / 256))."""
        props = torch.cuda.get_device_properties(self.device)
        blocks = min(props.multi_processor_count * (props.max_threads_per_multi_processor // 256), (numel + 255) // 256)
        return 256 * max(1, blocks)

    def row_absmax(self, x: Tensor) -> Tensor:
        x = self._f32(x)
        out = torch.empty(x.shape[0], dtype=torch.float32, device=self.device)
        check(self.lib.pdm_row_absmax_f32(x.data_ptr(), x.shape[0], x.shape[1], _ld(x), out.data_ptr(), self._stream()),
              "pdm_row_absmax_f32")
        self.launches += 1
        return out

    def noised_rows_philox(self, seed: int, offset: int, offset_step: int, x0: Tensor, sigma: Tensor, *,
                           x0_absmax: Optional[Tensor] = None, want_x: bool = False, want_split: bool = True) -> dict:
        """Rows (t, r) = fl(fl(randn * sigma[t]) + x0[r]) for the draws t at Philox offsets offset + t*offset_step;
        same keys as prepare_rows."""
        x0, sigma = self._f32(x0), self._f32(sigma)
        b, d = x0.shape
        n = sigma.shape[0]
        rows = n * b
        out = {"x": None, "norms": None, "hi": None, "lo": None, "inv_scale": None}
        if want_x:
            out["x"] = torch.empty(rows, d, dtype=torch.float32, device=self.device)
        if want_split:
            if x0_absmax is None:
                x0_absmax = self.row_absmax(x0)
            out["hi"] = torch.empty(rows, d, dtype=torch.float16, device=self.device)
            out["lo"] = torch.empty(rows, d, dtype=torch.float16, device=self.device)
            out["inv_scale"] = torch.empty(rows, dtype=torch.float32, device=self.device)
        check(self.lib.pdm_noised_rows_philox(
            seed, offset, offset_step, self.randn_launch_threads(b * d), x0.data_ptr(), b, d, _ld(x0), sigma.data_ptr(), n,
            _ptr(x0_absmax), _ptr(out["x"]), d, _ptr(out["hi"]), _ptr(out["lo"]), d, _ptr(out["inv_scale"]),
            self._stream()), "pdm_noised_rows_philox")
        self.launches += 1
        out["norms"] = torch.empty(rows, dtype=torch.float32, device=self.device)
        if want_split:
            check(self.lib.pdm_split_row_norms(out["hi"].data_ptr(), out["lo"].data_ptr(), d, out["inv_scale"].data_ptr(),
                                               rows, d, out["norms"].data_ptr(), self._stream()), "pdm_split_row_norms")
        else:
            check(self.lib.pdm_row_norms_f32(out["x"].data_ptr(), rows, d, d, out["norms"].data_ptr(), self._stream()),
                  "pdm_row_norms_f32")
        self.launches += 1
        return out

    # ---- row preparation -----------------------------------------------------------------------
    def prepare_rows(self, src: Tensor, rows: int, *, noise: Optional[Tensor] = None, sigma: Optional[Tensor] = None,
                     post: Optional[Tensor] = None, fixed_scale: float = 0.0, want_x: bool = False,
                     want_norms: bool = True, want_split: bool = True) -> dict:
        src = self._f32(src)
        d = src.shape[1]
        if noise is not None:
            noise, sigma = self._f32(noise), self._f32(sigma)
        out = {"x": None, "norms": None, "hi": None, "lo": None, "inv_scale": None}
        if want_x:
            out["x"] = torch.empty(rows, d, dtype=torch.float32, device=self.device)
        if want_norms:
            out["norms"] = torch.empty(rows, dtype=torch.float32, device=self.device)
        ldh = _round_up(d, 8)
        if want_split:
            out["hi"] = torch.empty(rows, ldh, dtype=torch.float16, device=self.device)
            out["lo"] = torch.empty(rows, ldh, dtype=torch.float16, device=self.device)
            out["inv_scale"] = torch.empty(rows, dtype=torch.float32, device=self.device)
        check(self.lib.pdm_prepare_rows(
            src.data_ptr(), src.shape[0], _ld(src),
            _ptr(noise), _ld(noise) if noise is not None else 0, _ptr(sigma), _ptr(post),
            rows, d, float(fixed_scale),
            _ptr(out["x"]), d, _ptr(out["norms"]),
            _ptr(out["hi"]), _ptr(out["lo"]), ldh, _ptr(out["inv_scale"]), self._stream()), "pdm_prepare_rows")
        self.launches += 1
        return out

    def transpose_split(self, y: Tensor, scale: float):
        y = self._f32(y)
        n, d = y.shape
        ldt = _round_up(n, 8)
        hi = torch.empty(d, ldt, dtype=torch.float16, device=self.device)
        lo = torch.empty(d, ldt, dtype=torch.float16, device=self.device)
        check(self.lib.pdm_transpose_split_f16(y.data_ptr(), n, d, _ld(y), float(scale), hi.data_ptr(),
                                               lo.data_ptr(), ldt, self._stream()), "pdm_transpose_split_f16")
        self.launches += 1
        return hi, lo

    def column_moments(self, y: Tensor):
        y = self._f32(y)
        n, d = y.shape
        s = torch.empty(d, dtype=torch.float64, device=self.device)
        s2 = torch.empty(d, dtype=torch.float64, device=self.device)
        mm = torch.empty(2, dtype=torch.float32, device=self.device)
        check(self.lib.pdm_column_moments_f32(y.data_ptr(), n, d, _ld(y), s.data_ptr(), s2.data_ptr(),
                                              mm.data_ptr(), self._stream()), "pdm_column_moments_f32")
        self.launches += 2
        return s, s2, mm

    # ---- fused pass ----------------------------------------------------------------------------
    def posterior_stats(self, *, precision: str, M: int, N: int, d: int, q_norm: Tensor, y_norm: Tensor,
                        inv_temp: Optional[Tensor], q: Optional[Tensor] = None, y: Optional[Tensor] = None,
                        q_split=None, y_split=None, y_inv_scale: float = 1.0, y_aux: Optional[Tensor] = None,
                        index_offset: int = 0, n_splits: int = 0, m_group: int = 0, cta_group: int = 0,
                        want_partials: bool = True, energy_out: Optional[Tensor] = None,
                        energy_mult: float = 1.0, row_tiles: Optional[Tensor] = None,
                        n_row_tiles: int = 0, n_row_tiles_dev: Optional[Tensor] = None, topk: bool = False,
                        plan_row_tiles: int = 0):
        """``row_tiles`` (int32, device) / ``n_row_tiles``: screened launch over the listed row tiles of
        128*cta_group rows only; records of the other rows are left as allocated (uninitialised).
        ``n_row_tiles_dev`` (one int32 on the device): the list's actual length, read by the kernel -- ``n_row_tiles`` is
        then only the upper bound, and nothing is read back to the host; ``plan_row_tiles`` is the length the schedule
        (m_group x n_splits) is planned for when the caller has an estimate (any value is safe: the kernel walks the
        device-side length whatever the plan)."""
        a = StatsArgs()
        a.precision = PRECISIONS[precision]
        a.n_splits, a.m_group, a.cta_group = n_splits, m_group, cta_group
        a.M, a.N, a.d, a.index_offset = M, N, d, index_offset
        keep = [q_norm, y_norm, inv_temp, y_aux, q, y, q_split, y_split, energy_out]
        if precision == "exact":
            q, y = self._f32(q), self._f32(y)
            a.q, a.ldq, a.y, a.ldy = q.data_ptr(), _ld(q), y.data_ptr(), _ld(y)
        else:
            q_hi, q_lo, q_inv = q_split
            y_hi, y_lo = y_split
            a.q_hi, a.q_lo, a.ldqh, a.q_inv_scale = q_hi.data_ptr(), _ptr(q_lo), _ld(q_hi), q_inv.data_ptr()
            a.y_hi, a.y_lo, a.ldyh, a.y_inv_scale = y_hi.data_ptr(), _ptr(y_lo), _ld(y_hi), float(y_inv_scale)
        a.q_norm, a.y_norm, a.inv_temp, a.y_aux = q_norm.data_ptr(), y_norm.data_ptr(), _ptr(inv_temp), _ptr(y_aux)
        if row_tiles is not None:
            assert row_tiles.dtype == torch.int32 and row_tiles.is_contiguous() and n_row_tiles <= row_tiles.numel()
            a.row_tiles, a.n_row_tiles = row_tiles.data_ptr(), int(n_row_tiles)
            keep.append(row_tiles)
            if n_row_tiles_dev is not None:
                assert n_row_tiles_dev.dtype == torch.int32
                a.n_row_tiles_dev = n_row_tiles_dev.data_ptr()
                keep.append(n_row_tiles_dev)
        nfloats = C.c_int64()
        if row_tiles is not None and plan_row_tiles > 0:
            a.n_row_tiles = max(1, min(int(plan_row_tiles), int(n_row_tiles)))
        check(self.lib.pdm_posterior_stats_plan(C.byref(a), self.device.index or 0, C.byref(nfloats)),
              "pdm_posterior_stats_plan")
        if row_tiles is not None:
            a.n_row_tiles = int(n_row_tiles)
        partials = None
        if topk:
            # top-k epilogue: per (record, row) the 8 smallest squared distances and their local dataset rows
            tk_val = torch.empty(a.records_per_row, M, _cabi.TOPK_SLOTS, dtype=torch.float32, device=self.device)
            tk_idx = torch.empty(a.records_per_row, M, _cabi.TOPK_SLOTS, dtype=torch.int32, device=self.device)
            a.topk_val, a.topk_idx = tk_val.data_ptr(), tk_idx.data_ptr()
            want_partials = False
        if want_partials:
            partials = torch.empty(a.records_per_row, M, _cabi.PART_STRIDE, dtype=torch.float32, device=self.device)
            a.partials = partials.data_ptr()
        if energy_out is not None:
            a.energy_out, a.lde, a.energy_mult = energy_out.data_ptr(), _ld(energy_out), float(energy_mult)
        if self.kernel_events is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        check(self.lib.pdm_posterior_stats(C.byref(a), self._stream()), "pdm_posterior_stats")
        if self.kernel_events is not None:
            ev1.record()
            rows_done = M if row_tiles is None else min(M, int(n_row_tiles) * 128 * a.cta_group)
            self.kernel_events.append((ev0, ev1, rows_done * N))
        self.launches += 1
        self.last_plan = (a.n_splits, a.m_group, a.cta_group)
        del keep
        if topk:
            return tk_val, tk_idx
        return partials

    def topk_merge(self, tk_val: Tensor, tk_idx: Tensor, index_offset: int, k: int):
        """records of the top-k epilogue -> (values (M, k) ascending, global indices (M, k) int64)."""
        records, m, _ = tk_val.shape
        vals = torch.empty(m, k, dtype=torch.float32, device=self.device)
        idx = torch.empty(m, k, dtype=torch.int64, device=self.device)
        check(self.lib.pdm_topk_merge(tk_val.data_ptr(), tk_idx.data_ptr(), m, records, int(index_offset), int(k),
                                      vals.data_ptr(), idx.data_ptr(), self._stream()), "pdm_topk_merge")
        self.launches += 1
        return vals, idx

    def merge(self, parts: Tensor, inv_temp: Tensor, n_total: int):
        """parts: (S, M, 8) record-major, as posterior_stats emits them, or (G, M, S, 8): G all-gathered shards of
        row-major records (the (M, 1, 8) outputs of ``reduce``)."""
        parts = parts.contiguous()
        if parts.dim() == 3:
            s, m, _ = parts.shape
            g, outer, inner, row = 1, 0, parts.stride(0), parts.stride(1)
        else:
            g, m, s, _ = parts.shape
            outer, inner, row = parts.stride(0), parts.stride(2), parts.stride(1)
        out = torch.empty(_cabi.OUT_ROWS, m, dtype=torch.float32, device=self.device)
        argmin = torch.empty(m, dtype=torch.int64, device=self.device)
        check(self.lib.pdm_merge_partials(parts.data_ptr(), m, g, outer, s, inner, row, inv_temp.data_ptr(), n_total,
                                          out.data_ptr(), argmin.data_ptr(), self._stream()), "pdm_merge_partials")
        self.launches += 1
        return out, argmin

    # ---- certified delta posteriors (adaptive precision) -------------------------------------------
    def screen_temperatures(self, q_norm: Tensor, inv_temp: Tensor, y_norm_max: Tensor, g: float, e_star: float,
                            kappa: float) -> Tensor:
        out = torch.empty_like(inv_temp)
        check(self.lib.pdm_screen_temperatures(q_norm.data_ptr(), inv_temp.data_ptr(), inv_temp.numel(), y_norm_max.data_ptr(),
                                               float(g), float(e_star), float(kappa), out.data_ptr(), self._stream()),
              "pdm_screen_temperatures")
        self.launches += 1
        return out

    def split_to_e4m3(self, hi: Tensor, lo: Optional[Tensor], d: int):
        """fp16 split operands -> (E4M3 bytes (rows, round_up(d, 16)), exact per-row rounding deviation in scaled units)."""
        rows = hi.shape[0]
        out8 = torch.empty(rows, _round_up(d, 16), dtype=torch.uint8, device=self.device)
        err = torch.empty(rows, dtype=torch.float32, device=self.device)
        check(self.lib.pdm_split_to_e4m3(hi.data_ptr(), _ptr(lo), _ld(hi), rows, d, out8.data_ptr(), _ld(out8), err.data_ptr(),
                                         self._stream()), "pdm_split_to_e4m3")
        self.launches += 1
        return out8, err

    def screen_temperatures_f8(self, q_norm: Tensor, q_err: Tensor, q_inv_scale: Tensor, inv_temp: Tensor, y_norm_max: Tensor,
                               y_err_max: Tensor, g: float, e_star: float, kappa: float) -> Tensor:
        out = torch.empty_like(inv_temp)
        check(self.lib.pdm_screen_temperatures_f8(q_norm.data_ptr(), q_err.data_ptr(), q_inv_scale.data_ptr(), inv_temp.data_ptr(),
                                                  inv_temp.numel(), y_norm_max.data_ptr(), y_err_max.data_ptr(), float(g),
                                                  float(e_star), float(kappa), out.data_ptr(), self._stream()),
              "pdm_screen_temperatures_f8")
        self.launches += 1
        return out

    def screen_tile_list(self, flags: Tensor, rows_per_tile: int):
        m = flags.numel()
        tile_list = torch.empty(max(1, (m + rows_per_tile - 1) // rows_per_tile), dtype=torch.int32, device=self.device)
        n_listed = torch.empty(1, dtype=torch.int32, device=self.device)
        check(self.lib.pdm_screen_tile_list(flags.data_ptr(), m, int(rows_per_tile), tile_list.data_ptr(), n_listed.data_ptr(),
                                            self._stream()), "pdm_screen_tile_list")
        self.launches += 1
        return tile_list, n_listed

    def screen_certify(self, screen_out: Tensor, e_star: float, rows_per_tile: int):
        """-> flags (M,) uint8, tile_list (tiles,) int32, n_listed (1,) int32 -- all on the device."""
        m = screen_out.shape[1]
        tiles = (m + rows_per_tile - 1) // rows_per_tile
        flags = torch.empty(m, dtype=torch.uint8, device=self.device)
        tile_list = torch.empty(max(1, tiles), dtype=torch.int32, device=self.device)
        n_listed = torch.empty(1, dtype=torch.int32, device=self.device)
        check(self.lib.pdm_screen_certify(screen_out.data_ptr(), m, float(e_star), int(rows_per_tile), flags.data_ptr(),
                                          tile_list.data_ptr(), n_listed.data_ptr(), self._stream()), "pdm_screen_certify")
        self.launches += 2
        return flags, tile_list, n_listed

    def screen_finalize(self, flags: Tensor, screen_argmin: Tensor, d: int, q_split, q_norm: Tensor, y_split,
                        y_inv_scale: float, y_norm: Tensor, y_aux: Optional[Tensor], index_offset: int, n_local: int,
                        n_total: int, out: Tensor, argmin: Tensor) -> None:
        q_hi, q_lo, q_inv = q_split
        y_hi, y_lo = y_split
        check(self.lib.pdm_screen_finalize(flags.data_ptr(), screen_argmin.data_ptr(), flags.numel(), d,
                                           q_hi.data_ptr(), _ptr(q_lo), _ld(q_hi), q_inv.data_ptr(), q_norm.data_ptr(),
                                           y_hi.data_ptr(), _ptr(y_lo), _ld(y_hi), float(y_inv_scale), y_norm.data_ptr(),
                                           _ptr(y_aux), index_offset, n_local, n_total, out.data_ptr(), argmin.data_ptr(),
                                           self._stream()), "pdm_screen_finalize")
        self.launches += 1

    def reduce(self, parts: Tensor, inv_temp: Tensor) -> Tensor:
        """(S, M, 8) record-major -> (M, 1, 8): one merged, not yet finalised, record per row."""
        parts = parts.contiguous()
        s, m, _ = parts.shape
        out = torch.empty(m, 1, _cabi.PART_STRIDE, dtype=torch.float32, device=self.device)
        check(self.lib.pdm_reduce_partials(parts.data_ptr(), m, 1, 0, s, parts.stride(0), parts.stride(1),
                                           inv_temp.data_ptr(), out.data_ptr(), self._stream()), "pdm_reduce_partials")
        self.launches += 1
        return out

    # ---- posterior mean ------------------------------------------------------------------------
    def weights_from_energy(self, energy: Tensor, e_min: Tensor, l: Tensor, inv_temp: Tensor, *, split: bool,
                            tiles=None):
        """``tiles`` = (row_tiles int32, rows_per_tile, upper bound of listed tiles, device count or None): only the rows
        of the listed tiles are formed; the other rows of the outputs stay as allocated."""
        m, n = energy.shape
        tl, rpt, n_max, n_dev = tiles if tiles is not None else (None, 0, 0, None)
        if split:
            ld = _round_up(n, 8)
            hi = torch.empty(m, ld, dtype=torch.float16, device=self.device)
            lo = torch.empty(m, ld, dtype=torch.float16, device=self.device)
            check(self.lib.pdm_weights_from_energy_tiles(energy.data_ptr(), _ld(energy), m, n, e_min.data_ptr(),
                                                         l.data_ptr(), inv_temp.data_ptr(), None, 0, hi.data_ptr(),
                                                         lo.data_ptr(), ld, _ptr(tl), int(rpt), int(n_max), _ptr(n_dev),
                                                         self._stream()), "pdm_weights_from_energy")
            self.launches += 1
            return hi, lo
        p = torch.empty(m, n, dtype=torch.float32, device=self.device)
        check(self.lib.pdm_weights_from_energy_tiles(energy.data_ptr(), _ld(energy), m, n, e_min.data_ptr(),
                                                     l.data_ptr(), inv_temp.data_ptr(), p.data_ptr(), n, None, None, 0,
                                                     _ptr(tl), int(rpt), int(n_max), _ptr(n_dev), self._stream()),
              "pdm_weights_from_energy")
        self.launches += 1
        return p

    def delta_tile_list(self, l: Tensor, rows_per_tile: int):
        """-> (flags uint8 (M,): posterior is a delta to fp32 resolution, tile_list int32, n_listed int32 (1,)): the row
        tiles that hold a row whose weights have to be contracted."""
        m = l.numel()
        flags = torch.empty(m, dtype=torch.uint8, device=self.device)
        tile_list = torch.empty(max(1, (m + rows_per_tile - 1) // rows_per_tile), dtype=torch.int32, device=self.device)
        n_listed = torch.empty(1, dtype=torch.int32, device=self.device)
        check(self.lib.pdm_delta_tile_list(l.data_ptr(), m, int(rows_per_tile), flags.data_ptr(), tile_list.data_ptr(),
                                           n_listed.data_ptr(), self._stream()), "pdm_delta_tile_list")
        self.launches += 2
        return flags, tile_list, n_listed

    def screen_merge_stage(self, tile_list: Tensor, n_listed: Tensor, max_tiles: int, rows_per_tile: int, flags_b: Tensor,
                           arg_b: Tensor, flags: Tensor, arg: Tensor) -> None:
        check(self.lib.pdm_screen_merge_stage(tile_list.data_ptr(), n_listed.data_ptr(), int(max_tiles), int(rows_per_tile),
                                              flags.numel(), flags_b.data_ptr(), arg_b.data_ptr(), flags.data_ptr(),
                                              arg.data_ptr(), self._stream()), "pdm_screen_merge_stage")
        self.launches += 1

    def gather_rows(self, src: Tensor, idx: Tensor, index_offset: int, flags: Optional[Tensor], out: Tensor) -> Tensor:
        """out[r] = src[idx[r] - index_offset] for flagged rows (zeros when another shard owns the index)."""
        src = self._f32(src)
        check(self.lib.pdm_gather_rows_f32(src.data_ptr(), _ld(src), src.shape[0], src.shape[1], idx.data_ptr(),
                                           int(index_offset), _ptr(flags), idx.numel(), out.data_ptr(), _ld(out),
                                           self._stream()), "pdm_gather_rows_f32")
        self.launches += 1
        return out

    def split_gemm(self, a_hi: Tensor, a_lo: Tensor, b_hi: Tensor, b_lo: Optional[Tensor], k: int, scale: float,
                   out: Optional[Tensor] = None, accumulate: bool = False, cta_group: int = 0, tiles=None) -> Tensor:
        """``tiles`` = (row_tiles int32, upper bound of listed tiles, device count or None): only the listed tiles of
        128*cta_group rows are contracted and written."""
        m, d = a_hi.shape[0], b_hi.shape[0]
        if out is None:
            out = torch.empty(m, d, dtype=torch.float32, device=self.device)
        tl, n_max, n_dev = tiles if tiles is not None else (None, 0, None)
        check(self.lib.pdm_split_gemm_f16x3_tiles(a_hi.data_ptr(), a_lo.data_ptr(), _ld(a_hi), m, b_hi.data_ptr(),
                                                  _ptr(b_lo), _ld(b_hi), d, k, float(scale), out.data_ptr(),
                                                  _ld(out), int(accumulate), cta_group, _ptr(tl), int(n_max), _ptr(n_dev),
                                                  self._stream()),
              "pdm_split_gemm_f16x3")
        self.launches += 1
        return out

    def weighted_mean_exact(self, p: Tensor, y: Tensor, out: Optional[Tensor] = None, accumulate: bool = False) -> Tensor:
        m, n = p.shape
        d = y.shape[1]
        if out is None:
            out = torch.empty(m, d, dtype=torch.float32, device=self.device)
        check(self.lib.pdm_weighted_mean_exact_f32(p.data_ptr(), _ld(p), m, n, y.data_ptr(), _ld(y), d,
                                                   out.data_ptr(), _ld(out), int(accumulate), self._stream()),
              "pdm_weighted_mean_exact_f32")
        self.launches += 1
        return out

    def denoiser_backward_weights(self, energy: Tensor, sdot: Tensor, e_min: Tensor, l: Tensor, inv_temp: Tensor,
                                  s_scale: Optional[Tensor] = None, a_in: Optional[Tensor] = None):
        """Centred weights w = p * (s - a) (M, N), s = s_scale * sdot, a = sum_j p_j s_j, and sums (M, 2) = (a, sum w e)."""
        m, n = energy.shape
        w = torch.empty(m, n, dtype=torch.float32, device=self.device)
        sums = torch.empty(m, 2, dtype=torch.float32, device=self.device)
        check(self.lib.pdm_denoiser_backward_weights(energy.data_ptr(), _ld(energy), sdot.data_ptr(), _ld(sdot),
                                                     m, n, e_min.data_ptr(), l.data_ptr(), inv_temp.data_ptr(),
                                                     _ptr(s_scale), _ptr(a_in), w.data_ptr(), _ld(w), sums.data_ptr(),
                                                     self._stream()), "pdm_denoiser_backward_weights")
        self.launches += 1
        return w, sums

    def topk_smallest(self, x: Tensor, k: int):
        """k smallest entries per row of a dense (rows, n) matrix: (values (rows, k) ascending, indices int64)."""
        x = self._f32(x)
        rows, n = x.shape
        vals = torch.empty(rows, k, dtype=torch.float32, device=self.device)
        idx = torch.empty(rows, k, dtype=torch.int64, device=self.device)
        check(self.lib.pdm_topk_smallest_f32(x.data_ptr(), _ld(x), rows, n, int(k), vals.data_ptr(), idx.data_ptr(),
                                             self._stream()), "pdm_topk_smallest_f32")
        self.launches += 1
        return vals, idx

    def refine_neighbours(self, x: Tensor, y: Tensor, index_offset: int, vals: Tensor, idx: Tensor) -> None:
        """In place: exact squared distances (direct differences, fp64 accumulation) of the candidates, rows re-sorted."""
        x, y = self._f32(x), self._f32(y)
        check(self.lib.pdm_refine_neighbours_f32(x.data_ptr(), _ld(x), x.shape[0], x.shape[1], y.data_ptr(), _ld(y), y.shape[0],
                                                 int(index_offset), vals.shape[1], vals.data_ptr(), idx.data_ptr(),
                                                 self._stream()), "pdm_refine_neighbours_f32")
        self.launches += 1

    def sampler_step(self, x0_hat: Tensor, xt: Tensor, noise: Optional[Tensor], c_x0: float, c_xt: float, c_noise: float,
                     out: Optional[Tensor] = None, coef: Optional[Tensor] = None) -> Tensor:
        """out = c_x0 * x0_hat + c_xt * xt (+ c_noise * noise); ``out`` may be ``xt``.  ``coef`` (3 floats on the device)
        replaces the three host scalars (CUDA-graphed sampling steps)."""
        if out is None:
            out = torch.empty_like(xt)
        for t in (x0_hat, xt, out) + ((noise,) if noise is not None else ()):
            if t.dtype != torch.float32 or not t.is_contiguous() or t.device != self.device:
                raise PdmError("sampler_step needs contiguous float32 tensors on the engine's device")
        if coef is not None:
            check(self.lib.pdm_sampler_step_dev_f32(x0_hat.data_ptr(), xt.data_ptr(), _ptr(noise), coef.data_ptr(),
                                                    out.data_ptr(), xt.numel(), self._stream()), "pdm_sampler_step_dev_f32")
            self.launches += 1
            return out
        check(self.lib.pdm_sampler_step_f32(x0_hat.data_ptr(), xt.data_ptr(), _ptr(noise), float(c_x0), float(c_xt),
                                            float(c_noise), out.data_ptr(), xt.numel(), self._stream()), "pdm_sampler_step_f32")
        self.launches += 1
        return out
