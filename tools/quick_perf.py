"""Quick device-time probe of the fused tensor kernel on a CIFAR-10-shaped block (dev tool, not the bench)."""
import argparse
import os
import sys
import time

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
    ap.add_argument("--configs", type=str, default="2:0:0,1:0:0")   # cta_group:m_group:n_splits
    ap.add_argument("--precision", type=str, default="f16x3")
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args()
    be = CudaBackend()
    dev = be.device
    torch.manual_seed(0)
    y = torch.rand(a.n, a.d, device=dev) * 2 - 1
    x = y[torch.randint(0, a.n, (a.m,), device=dev)] + 0.3 * torch.randn(a.m, a.d, device=dev)
    inv_t = torch.full((a.m,), 1.0 / 0.09, device=dev)
    y_norm = be.row_norms(y)
    scale = pow2_scale_for(float(be.absmax(y).item()))
    ys = be.prepare_rows(y, a.n, fixed_scale=scale, want_norms=False)
    t0 = time.time()
    prep = be.prepare_rows(x, a.m)
    torch.cuda.synchronize()
    print(f"prepare_rows({a.m}x{a.d}) first call {1e3 * (time.time() - t0):.2f} ms", flush=True)
    for cfg in a.configs.split(","):
        cg, mg, ns = (int(v) for v in cfg.split(":"))
        def run():
            return be.posterior_stats(precision=a.precision, M=a.m, N=a.n, d=a.d, q_norm=prep["norms"], y_norm=y_norm,
                                      inv_temp=inv_t, q_split=(prep["hi"], prep["lo"], prep["inv_scale"]),
                                      y_split=(ys["hi"], ys["lo"]), y_inv_scale=1.0 / scale, cta_group=cg, m_group=mg,
                                      n_splits=ns)
        run()
        torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters + 1)]
        evs[0].record()
        for i in range(a.iters):
            parts = run()
            evs[i + 1].record()
        torch.cuda.synchronize()
        each = [evs[i].elapsed_time(evs[i + 1]) for i in range(a.iters)]
        ms = sum(each) / a.iters
        if a.iters <= 12:
            print("   per-iter ms:", " ".join(f"{v:.2f}" for v in each), flush=True)
        else:
            tail = each[len(each) // 2:]
            print(f"   first {each[0]:.2f}  min {min(each):.2f}  steady(mean of last half) {sum(tail) / len(tail):.2f} ms", flush=True)
            ms = sum(tail) / len(tail)
        pairs = a.m * a.n
        terms = 3 if a.precision == "f16x3" else 1
        print(f"cfg cg={cg} plan(S,G,cg)={be.last_plan}: {ms:.3f} ms  {pairs / ms / 1e6:.2f} Gpairs/s  "
              f"algorithmic {2 * a.d * pairs / ms / 1e9:.1f} TFLOP/s  executed {terms * 2 * a.d * pairs / ms / 1e9:.1f} TFLOP/s",
              flush=True)


if __name__ == "__main__":
    main()
