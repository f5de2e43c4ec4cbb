"""Host-side (Python) profile of one noised_stats call at the size of one rank of the 8-GPU grid (dev tool)."""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "physics-of-diffusion-models_b200"))
sys.path.insert(0, ROOT)
from pdm_b200 import EmpiricalDataset, PosteriorEngine, EngineConfig  # noqa: E402
from pdm_b200.backend import CudaBackend  # noqa: E402
from bench import ddpm_temperatures  # noqa: E402

n, d, b = int(os.environ.get("N", 25000)), int(os.environ.get("D", 3072)), int(os.environ.get("B", 1024))
be = CudaBackend()
dev = be.device
torch.manual_seed(0)
y = torch.rand(n, d, device=dev) * 2 - 1
eng = PosteriorEngine(EmpiricalDataset(y, backend=be), EngineConfig())
temps = ddpm_temperatures(1000, 1e-4, 2.478e4).to(dev)[0::int(os.environ.get("STRIDE", 4))].contiguous()
x0 = y[:b].clone()
for i in range(6):                      # the first calls one by one: probing call, first remembered-boundary call, steady state
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.noised_stats(x0, temps)
    torch.cuda.synchronize()
    print(f"call {i}: {(time.perf_counter() - t0) * 1e3:.2f} ms, reserved {torch.cuda.memory_reserved() / 2**30:.2f} GiB, "
          f"boundary {eng._screen_prior}")
t0 = time.perf_counter()
for _ in range(5):
    eng.noised_stats(x0, temps)
torch.cuda.synchronize()
print(f"wall per call, 5 back-to-back calls: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms")
pr = cProfile.Profile()
pr.enable()
eng.noised_stats(x0, temps)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
