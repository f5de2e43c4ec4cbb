"""Dev probe (needs lib built with -DPDM_STALL_STATS, loaded through PDM_B200_LIB): per-launch time of the
fused kernel next to the time its roles spent blocked on their barriers."""
import argparse
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "physics-of-diffusion-models_b200"))

from pdm_b200.backend import CudaBackend  # noqa: E402
from pdm_b200.engine import pow2_scale_for  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=16384)
    ap.add_argument("--n", type=int, default=50000)
    ap.add_argument("--d", type=int, default=3072)
    ap.add_argument("--iters", type=int, default=12)
    ap.add_argument("--precs", type=str, default="f16x3,f16x2,f16x1")
    a = ap.parse_args()
    be = CudaBackend()
    lib = be.lib
    mc = C.c_int()
    lib.pdm_debug_max_clusters(C.byref(mc))
    print("max co-resident 2-CTA clusters:", mc.value, " SMs:", be.sm_count, flush=True)
    dev = be.device
    torch.manual_seed(0)
    y = torch.rand(a.n, a.d, device=dev) * 2 - 1
    x = y[torch.randint(0, a.n, (a.m,), device=dev)] + 0.3 * torch.randn(a.m, a.d, device=dev)
    inv_t = torch.full((a.m,), 1.0 / 0.09, device=dev)
    y_norm = be.row_norms(y)
    scale = pow2_scale_for(float(be.absmax(y).item()))
    ys = be.prepare_rows(y, a.n, fixed_scale=scale, want_norms=False)
    prep = be.prepare_rows(x, a.m)
    buf = (C.c_ulonglong * (160 * 8))()
    for prec in a.precs.split(","):
        print(f"--- {prec}   (ms | mean over CTAs of blocked ms: producer<-empty, mma<-full, mma<-tmem_empty, epi<-tmem_full | max wall)")
        for it in range(a.iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            be.posterior_stats(precision=prec, M=a.m, N=a.n, d=a.d, q_norm=prep["norms"], y_norm=y_norm, inv_temp=inv_t,
                               q_split=(prep["hi"], prep["lo"], prep["inv_scale"]), y_split=(ys["hi"], ys["lo"]),
                               y_inv_scale=1.0 / scale, cta_group=2)
            e1.record()
            torch.cuda.synchronize()
            lib.pdm_debug_read_stalls(buf)
            t = torch.tensor(list(buf), dtype=torch.float64).view(160, 8)[:148] / 1e6
            lead = t[0::2]          # leader CTAs carry the MMA counters
            print(f"  {e0.elapsed_time(e1):8.3f} | {t[:, 0].mean():7.3f} {lead[:, 1].mean():7.3f} {lead[:, 2].mean():7.3f} "
                  f"{t[:, 3].mean():7.3f} | {t[:, 4].max():7.3f} | stats phase {t[:, 5].mean():6.3f}   (max over CTAs: {t[:, 0].max():.2f} {lead[:, 1].max():.2f} "
                  f"{lead[:, 2].max():.2f} {t[:, 3].max():.2f})", flush=True)


if __name__ == "__main__":
    main()
