"""Dev probe: error of the tensor-core squared distances against fp64, next to the reference's own fp32 error.
Env knobs are read by the library (PDM_FLUSH_KB, PDM_B200_LIB)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "physics-of-diffusion-models_b200"))
sys.path.insert(0, ROOT)

from pdm_b200.backend import CudaBackend  # noqa: E402
from pdm_b200.engine import pow2_scale_for  # noqa: E402


def sqdist(x, y):
    xn = (x * x).sum(1)
    yn = (y * y).sum(1)
    return xn[:, None] - 2 * (x @ y.t()) + yn


def main():
    be = CudaBackend()
    dev = be.device
    g = torch.Generator().manual_seed(9)
    n, d, m = 2048, 3072, 256
    for kind in ("uniform", "lattice"):
        if kind == "uniform":
            data = torch.rand(n, d, generator=g) * 2 - 1
        else:
            px = torch.randint(0, 256, (n, d), generator=g, dtype=torch.uint8)
            data = (px.float() / 255 - 0.5) / 0.5
        for sigma in (0.05, 1.0, 30.0):
            xq = data[:m] + sigma * torch.randn(m, d, generator=g)
            y = data.to(dev)
            y_norm = be.row_norms(y)
            ref64 = sqdist(xq.double(), data.double())
            ref32 = sqdist(xq, data).double()
            e32 = ref32 - ref64
            ulp = 2.0 ** -23 * ((xq.double() ** 2).sum(1)[:, None] + (data.double() ** 2).sum(1))   # ~ulp of the norms
            print(f"{kind:8s} sigma={sigma:5.2f}  reference fp32: max {e32.abs().max():.2e} rms {e32.pow(2).mean().sqrt():.2e}"
                  f"  (in norm-ulps: max {(e32.abs() / ulp).max():.2f})")
            precs = [("f16x3", pow2_scale_for(float(be.absmax(y).item())))]
            if kind == "lattice":
                precs.append(("f16x2", 255.0))
            for prec, scale in precs:
                ys = be.prepare_rows(y, n, fixed_scale=scale, want_norms=False)
                prep = be.prepare_rows(xq.to(dev), m)
                dist = torch.empty(m, n, device=dev)
                be.posterior_stats(precision=prec, M=m, N=n, d=d, q_norm=prep["norms"], y_norm=y_norm, inv_temp=None,
                                   q_split=(prep["hi"], prep["lo"], prep["inv_scale"]), y_split=(ys["hi"], ys["lo"]),
                                   y_inv_scale=1.0 / scale, want_partials=False, energy_out=dist, energy_mult=2.0)
                err = dist.cpu().double() - ref64
                print(f"      {prec} scale={scale:g}: max {err.abs().max():.2e} mean {err.mean():+.2e} rms {err.pow(2).mean().sqrt():.2e}"
                      f"  (in norm-ulps: max {(err.abs() / ulp).max():.2f})   FLUSH_KB={os.environ.get('PDM_FLUSH_KB', '1')}")


if __name__ == "__main__":
    main()
