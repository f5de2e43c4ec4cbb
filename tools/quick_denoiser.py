"""Device-time probe of the posterior-mean (ideal denoiser) path at CIFAR-10 shape (dev tool)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "physics-of-diffusion-models_b200"))
from pdm_b200 import EmpiricalDataset, PosteriorEngine, EngineConfig
from pdm_b200.backend import CudaBackend

n, d, b = int(os.environ.get("N", 50000)), int(os.environ.get("D", 3072)), int(os.environ.get("B", 10000))
be = CudaBackend()
torch.manual_seed(0)
y = torch.rand(n, d, device=be.device) * 2 - 1
ds = EmpiricalDataset(y, backend=be)
eng = PosteriorEngine(ds, EngineConfig())
ab = torch.tensor(float(os.environ.get("AB", 0.002)), device=be.device)        # T ~ 500: every point carries weight
x = ab.sqrt() * y[torch.randint(0, n, (b,), device=be.device)] + (1 - ab).sqrt() * torch.randn(b, d, device=be.device)
t = ((1 - ab) / ab).expand(b)
post = ab.rsqrt().expand(b)
eng.posterior_mean(x, t, post=post)
torch.cuda.synchronize()
be.phase_events = {}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    out = eng.posterior_mean(x, t, post=post)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"posterior_mean B={b} N={n} d={d}: {ms:.2f} ms/step  {b * n / ms / 1e6:.2f} Gpairs/s  "
      f"algorithmic(4d) {4 * d * b * n / ms / 1e9:.1f} TFLOP/s", flush=True)
print("phases ms/call:", {k: round(v / 3, 2) for k, v in be.phase_totals().items()})
