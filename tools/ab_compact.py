"""A/B of EngineConfig.screen_compact on the C2 workload in ONE process (same box, same clocks), interleaved (dev tool)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "physics-of-diffusion-models_b200"))
sys.path.insert(0, ROOT)
from pdm_b200 import EmpiricalDataset, PosteriorEngine, EngineConfig  # noqa: E402
from pdm_b200.backend import CudaBackend  # noqa: E402
from bench import ddpm_temperatures  # noqa: E402

n, d, b = int(os.environ.get("N", 50000)), 3072, 1024
be = CudaBackend()
dev = be.device
torch.manual_seed(0)
kind = os.environ.get("DATA", "uniform")
if kind == "pixels":
    y = (torch.randint(0, 256, (n, d), device=dev, dtype=torch.uint8).float() / 255 - 0.5) / 0.5
else:
    y = torch.rand(n, d, device=dev) * 2 - 1
ds = EmpiricalDataset(y, backend=be)
temps = ddpm_temperatures(1000, 1e-4, 2.478e4).to(dev)
x0 = y[:b].clone()
engines = {c: PosteriorEngine(ds, EngineConfig(screen_compact=c)) for c in (True, False)}
for e in engines.values():
    for _ in range(3):
        e.noised_stats(x0, temps)
ref = None
for rnd in range(3):
    for c, e in engines.items():
        torch.cuda.synchronize()
        rep0 = dict(e.screen_report)
        t0 = time.perf_counter()
        for i in range(3):
            torch.manual_seed(100 + i)
            out = e.noised_stats(x0, temps)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 3 * 1e3
        full = (e.screen_report["tiles_full_pass"] - rep0["tiles_full_pass"]) // 3
        print(f"round {rnd} compact={c}: {ms:.2f} ms per call, full-pass tiles among screened rows {full}", flush=True)
        if ref is None:
            ref = out
        else:
            same = all(torch.equal(out[k], ref[k]) for k in ("entropy", "log_l", "var_e", "e_min", "argmin"))
            print(f"   identical to the first engine's results: {same}", flush=True)
