"""Dev probe: per-step time of IdealSampler on trajectory slices of the C5 shape, eager loop against CUDA-graphed steps."""
import math
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "physics-of-diffusion-models_b200"))

from bench import ddpm_temperatures  # noqa: E402
from pdm_b200 import EmpiricalDataset, EngineConfig, IdealSampler, PosteriorEngine  # noqa: E402
from pdm_b200.backend import CudaBackend  # noqa: E402


def main():
    be = CudaBackend()
    dev = be.device
    n, d, m = 50_000, 3072, int(os.environ.get("M", 10_000))
    torch.manual_seed(0)
    data = torch.rand(n, d, device=dev) * 2 - 1
    log_t = ddpm_temperatures(1000, 1e-4, 2.478e4).log().double()
    ds = EmpiricalDataset(data, backend=be)
    for i0, i1 in ((100, 150), (550, 600)):
        for graphs in (False, True):
            eng = PosteriorEngine(ds, EngineConfig(screen=True))
            smp = IdealSampler(data.view(n, 3, 32, 32), log_t[i0:i1], step_type="ddpm", engine=eng, use_graphs=graphs)
            ab = torch.sigmoid(-log_t[i1 - 1]).float().to(dev)
            torch.manual_seed(3)
            x_init = (ab.sqrt() * data[torch.randint(0, n, (m,), device=dev)] + (1 - ab).sqrt() * torch.randn(m, d, device=dev)).view(m, 3, 32, 32)
            outs = []
            for rep in range(3):
                torch.manual_seed(5)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = smp.batch_sample(m, x_init=x_init.clone())["x"]
                e1.record()
                t_host = time.perf_counter() - t0
                torch.cuda.synchronize()
                outs.append(out)
                print(f"steps {i1}->{i0} graphs={graphs} pass {rep}: {e0.elapsed_time(e1) / (i1 - i0):.3f} ms/step on the device, "
                      f"host enqueue {1e3 * t_host / (i1 - i0):.3f} ms/step, replays {smp.graph_replays}, "
                      f"marks screen_t={eng._pm_screen_t:.3g} f8_t={eng._pm_f8_t:.3g}, keys {len(smp._graph_seen)}", flush=True)
            print("   report:", {k: v for k, v in eng.screen_report.items() if k.startswith("pm_")},
                  "max diff between passes", (outs[1] - outs[2]).abs().max().item())


if __name__ == "__main__":
    main()
