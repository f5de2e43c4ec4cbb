"""Dev probe: GB/s of the HBM-bound kernels at the C2 shapes (row norms, dataset split, column moments, lattice test,
merge / reduce of partial records, weights, split-operand norms)."""
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "physics-of-diffusion-models_b200"))
from pdm_b200.backend import CudaBackend  # noqa: E402

be = CudaBackend()
dev = be.device
n, d = 50_000, 3072
y = torch.rand(n, d, device=dev) * 2 - 1
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=5):
    fn()
    ms = []
    for _ in range(reps):
        flush.zero_()                                   # evict L2 (126 MB) between repetitions
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return sorted(ms)[len(ms) // 2]


def report(name, nbytes, fn):
    ms = timed(fn)
    print(f"{name:34s} {nbytes / 1e9:7.3f} GB  {ms:8.3f} ms  {nbytes / ms / 1e6:8.1f} GB/s  ({nbytes / ms / 1e6 / 6552.6:5.1%} of 6552.6)")


report("row_norms (N x d fp32 read)", n * d * 4, lambda: be.row_norms(y))
report("absmax", n * d * 4, lambda: be.absmax(y))
report("prepare_rows dataset split (r4 w4)", n * d * 8, lambda: be.prepare_rows(y, n, fixed_scale=4096.0, want_norms=False))
report("column_moments", n * d * 4, lambda: be.column_moments(y))
report("lattice_residual", n * d * 4, lambda: be.lattice_residual(y, 2040.0))
report("transpose_split (r4 w4)", n * d * 8, lambda: be.transpose_split(y, 4096.0))
m, s = 172032, 12
parts = torch.rand(s, m, 8, device=dev)          # record-major, as the fused kernel writes them
parts[..., 1] += 0.5
inv_t = torch.ones(m, device=dev)
report("merge_partials (M x 12 records)", m * s * 32 + m * 40, lambda: be.merge(parts, inv_t, n))
report("reduce_partials", m * s * 32 + m * 32, lambda: be.reduce(parts, inv_t))
rows = 8053
energy = torch.rand(rows, n, device=dev)
e_min = torch.zeros(rows, device=dev)
l = torch.ones(rows, device=dev)
it = torch.ones(rows, device=dev)
report("weights_from_energy (r4 w4)", rows * n * 8, lambda: be.weights_from_energy(energy, e_min, l, it, split=True))
hi = torch.zeros(m // 4, d, dtype=torch.float16, device=dev)
lo = torch.zeros_like(hi)
inv = torch.ones(m // 4, device=dev)
nrm = torch.empty(m // 4, device=dev)
import ctypes  # noqa: E402
report("split_row_norms (r4)", (m // 4) * d * 4, lambda: be.lib.pdm_split_row_norms(hi.data_ptr(), lo.data_ptr(), d, inv.data_ptr(), m // 4, d, nrm.data_ptr(), be._stream()))
x0 = torch.rand(1024, d, device=dev)
st = torch.rand(8053, n, device=dev)
report("topk_smallest k=6 (6 row passes)", 8053 * n * 4, lambda: be.topk_smallest(st, 6))
report("sampler_step (r8 w4)", 10000 * d * 12, lambda: be.sampler_step(torch.empty(10000, d, device=dev), torch.empty(10000, d, device=dev), None, 0.5, 0.5, 0.0))
