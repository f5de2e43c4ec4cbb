"""Dev probe: config C1 (anisotropic GMM, N=10k, d=64, 100 temperatures, B=1024) through the engine, exact vs tensor path."""
import os
import sys
import time
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "physics-of-diffusion-models_b200"))
sys.path.insert(0, ROOT)
from oracle import synthetic as syn  # noqa: E402
from oracle import posterior as orc  # noqa: E402
from pdm_b200 import EmpiricalDataset, PosteriorEngine, EngineConfig  # noqa: E402
from pdm_b200.backend import CudaBackend  # noqa: E402

be = CudaBackend()
for dim in (64, 128, 192):
    data = syn.anisotropic_gmm(dim, 5, 10_000, 42)
    x0 = data[:1024].clone()
    temp = torch.logspace(-4, 4, 100)
    ds = EmpiricalDataset(data, backend=be)
    res = {}
    for prec in ("exact", "f16x3"):
        eng = PosteriorEngine(ds, EngineConfig(precision=prec))
        torch.manual_seed(0)
        eng.noised_stats(x0, temp)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            torch.manual_seed(0)
            st = eng.noised_stats(x0, temp)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        res[prec] = st["entropy"].cpu()
        print(f"d={dim} {prec}: {1e3 * dt:.2f} ms per batch  ({1024 * 100 * 10000 / dt / 1e9:.1f} Gpairs/s)")
    # fp64 oracle on the same noised queries (regenerate with torch on the device: same stream)
    torch.manual_seed(0)
    xt = torch.stack([torch.randn(1024, dim, device=be.device) * t.sqrt() + x0.to(be.device) for t in temp.to(be.device)]).cpu()
    ref = orc.entropy_batch(xt[::10], data, temp[::10], dtype=torch.float64)
    for prec in res:
        err = (res[prec][::10].double() - ref).abs()
        print(f"   {prec}: entropy |err| vs fp64: max {err.max():.2e}, per-temperature max {[f'{v:.1e}' for v in err.max(1).values.tolist()]}")
