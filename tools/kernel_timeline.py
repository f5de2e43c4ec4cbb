"""Per-kernel device time of one call of the hot paths (dev tool; torch.profiler = CUPTI, sees the library's kernels too).
  python tools/kernel_timeline.py denoise_low | denoise_mid | stats_low | stats_all"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "physics-of-diffusion-models_b200"))
sys.path.insert(0, ROOT)
from pdm_b200 import EmpiricalDataset, PosteriorEngine, EngineConfig  # noqa: E402
from pdm_b200.backend import CudaBackend  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "denoise_low"
n, d = int(os.environ.get("N", 50000)), int(os.environ.get("D", 3072))
be = CudaBackend()
dev = be.device
torch.manual_seed(0)
y = torch.rand(n, d, device=dev) * 2 - 1
eng = PosteriorEngine(EmpiricalDataset(y, backend=be), EngineConfig())

if what.startswith("denoise"):
    b = int(os.environ.get("B", 10000))
    ab = torch.tensor({"denoise_low": 0.5, "denoise_mid": 0.03}.get(what, 0.002), device=dev)
    x = ab.sqrt() * y[torch.randint(0, n, (b,), device=dev)] + (1 - ab).sqrt() * torch.randn(b, d, device=dev)
    t = ((1 - ab) / ab).expand(b)
    post = ab.rsqrt().expand(b)
    tb = (float(t[0]), float(t[0]))

    def call():
        return eng.posterior_mean(x, t, post=post, temp_bounds=tb)
else:
    b = int(os.environ.get("B", 1024))
    from bench import ddpm_temperatures  # noqa: E402
    temps = ddpm_temperatures(1000, 1e-4, 2.478e4).to(dev)
    if what == "stats_low":
        temps = temps[:168]
    if what == "stats_rank8":             # the share of one rank of the 8-GPU grid (dataset rows / 2 x temperatures / 4), run with N=25000
        temps = temps[0::4].contiguous()
    x0 = y[:b].clone()

    def call():
        return eng.noised_stats(x0, temps)["entropy"]

for _ in range(3):
    call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
call()
e1.record()
torch.cuda.synchronize()
print(f"{what}: {e0.elapsed_time(e1):.3f} ms per call (events, no profiler)")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    call()
    torch.cuda.synchronize()
rows = {}
total = 0.0
for ev in prof.events():
    if ev.device_type.name != "CUDA":
        continue
    us = ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    r = rows.setdefault(ev.name[:110], [0, 0.0])
    r[0] += 1
    r[1] += us
    total += us
print(f"device time under the profiler: {total / 1e3:.3f} ms in {sum(r[0] for r in rows.values())} kernels / copies")
for name, (cnt, us) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{us / 1e3:9.3f} ms {100 * us / total:6.2f} % {cnt:5d} x  {name}")

# where the device idles: gaps between consecutive kernels on the timeline (largest first)
evs = sorted(((ev.time_range.start, ev.time_range.end, ev.name[:60]) for ev in prof.events() if ev.device_type.name == "CUDA"),
             key=lambda t: t[0])
gaps = []
for (s0, e0_, n0), (s1, e1_, n1) in zip(evs, evs[1:]):
    if s1 > e0_:
        gaps.append((s1 - e0_, n0, n1))
span = (evs[-1][1] - evs[0][0]) if evs else 0
print(f"timeline span {span / 1e3:.3f} ms, idle {sum(g[0] for g in gaps) / 1e3:.3f} ms in {len(gaps)} gaps; largest:")
for g, n0, n1 in sorted(gaps, key=lambda t: -t[0])[:14]:
    print(f"{g / 1e3:8.3f} ms  after {n0}  before {n1}")
