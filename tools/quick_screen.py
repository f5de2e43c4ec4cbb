"""Dev probe: PosteriorEngine.noised_stats with and without certified delta posteriors (EngineConfig.screen) on a
synthetic shape, e.g. config C4:  python tools/quick_screen.py --n 200000 --d 12288 --b 1024 --nt 100 --tmin 1e-4 --tmax 1e8"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "physics-of-diffusion-models_b200"))

from pdm_b200 import EmpiricalDataset, EngineConfig, PosteriorEngine  # noqa: E402
from pdm_b200.backend import CudaBackend  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=50000)
    ap.add_argument("--d", type=int, default=3072)
    ap.add_argument("--b", type=int, default=1024)
    ap.add_argument("--nt", type=int, default=100)
    ap.add_argument("--tmin", type=float, default=1e-4)
    ap.add_argument("--tmax", type=float, default=1e4)
    ap.add_argument("--data", default="uniform", choices=["uniform", "sphere", "pixels"])
    ap.add_argument("--iters", type=int, default=2)
    a = ap.parse_args()
    be = CudaBackend()
    dev = be.device
    torch.manual_seed(0)
    if a.data == "uniform":
        data = torch.rand(a.n, a.d, device=dev) * 2 - 1
    elif a.data == "sphere":                                  # utils/synthetic_datasets.py:14-17, radius sqrt(d)
        data = torch.randn(a.n, a.d, device=dev)
        data *= (a.d ** 0.5) / data.norm(dim=1, keepdim=True)
    else:
        data = (torch.randint(0, 256, (a.n, a.d), device=dev, dtype=torch.uint8).float() / 255 - 0.5) / 0.5
    ds = EmpiricalDataset(data, backend=be)
    x0 = data[:a.b].clone()
    temps = torch.logspace(torch.log10(torch.tensor(a.tmin)).item(), torch.log10(torch.tensor(a.tmax)).item(), a.nt, device=dev)
    res = {}
    for screen in (False, True):
        eng = PosteriorEngine(ds, EngineConfig(screen=screen))
        torch.manual_seed(1)
        res[screen] = eng.noised_stats(x0, temps)            # warm-up (dataset split, plans)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.iters):
            torch.manual_seed(1)
            res[screen] = eng.noised_stats(x0, temps)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        pairs = a.b * a.nt * a.n
        print(f"screen={screen}: precision {eng.precision()}  {ms:.1f} ms/step  {pairs / ms / 1e6:.1f} G pairs/s  "
              f"algorithmic {2 * a.d * pairs / ms / 1e9:.0f} TFLOP/s  report {eng.screen_report if screen else ''}", flush=True)
    for k in ("entropy", "log_l", "mean_e"):
        print(f"  max |screened - unscreened| {k}: {(res[True][k] - res[False][k]).abs().max().item():.3e}")
    print("  argmin equal:", bool(torch.equal(res[True]["argmin"], res[False]["argmin"])))


if __name__ == "__main__":
    main()
