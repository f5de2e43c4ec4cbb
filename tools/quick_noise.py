"""Dev probe: in-kernel Philox noising + split vs torch.randn + pdm_prepare_rows, one block of the C2 step."""
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "physics-of-diffusion-models_b200"))
from pdm_b200.backend import CudaBackend  # noqa: E402
from pdm_b200 import PosteriorEngine  # noqa: E402

be = CudaBackend()
dev = be.device
b, d, nb = 1024, 3072, 56
x0 = torch.rand(b, d, device=dev) * 2 - 1
temps = torch.logspace(-3, 3, nb, device=dev)
sig = temps.sqrt().contiguous()
amax = be.row_absmax(x0)
step = PosteriorEngine._randn_offset_step((b, d), dev)


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def unfused():
    noise = torch.empty(nb, b, d, device=dev)
    for i in range(nb):
        torch.randn(b, d, device=dev, out=noise[i])
    return be.prepare_rows(x0, nb * b, noise=noise.view(nb * b, d), sigma=temps.repeat_interleave(b).sqrt())


print(f"torch.randn x{nb} + prepare_rows : {timed(unfused):.3f} ms")
print(f"noised_rows_philox (split+norms)  : {timed(lambda: be.noised_rows_philox(1, 0, step, x0, sig, x0_absmax=amax)):.3f} ms")
print(f"noised_rows_philox (fp32 x only)  : {timed(lambda: be.noised_rows_philox(1, 0, step, x0, sig, want_x=True, want_split=False)):.3f} ms")
gb = nb * b * d * 4 / 1e9
print(f"block = {gb:.2f} GB of fp32 noise / of fp16 hi+lo")
