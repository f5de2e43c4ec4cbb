"""Dev probe: how much of the fused kernel's power-capped rate depends on the operand bit patterns.

Runs the same f16x3 launch on a CIFAR-10-shaped block with the fp16 operands masked in different ways
(the arithmetic result is irrelevant here; only the steady-state rate under the power cap is read)."""
import argparse
import os
import subprocess
import sys
import threading

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "physics-of-diffusion-models_b200"))

from pdm_b200.backend import CudaBackend  # noqa: E402
from pdm_b200.engine import pow2_scale_for  # noqa: E402


class Smi:
    def __init__(self):
        self.rows = []
        self.proc = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw",
                                      "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
        threading.Thread(target=self._read, daemon=True).start()

    def _read(self):
        for line in self.proc.stdout:
            try:
                c, p = (float(v) for v in line.split(","))
                self.rows.append((c, p))
            except ValueError:
                pass

    def mark(self):
        return len(self.rows)

    def since(self, k):
        r = self.rows[k:]
        r = r[len(r) // 2:]
        if not r:
            return (0.0, 0.0)
        return (sum(c for c, _ in r) / len(r), sum(p for _, p in r) / len(r))


def mask16(t, keep_explicit_bits):
    """Round-to-zero an fp16 tensor to `keep_explicit_bits` explicit mantissa bits (10 = unchanged)."""
    if keep_explicit_bits >= 10:
        return t.clone()
    m = (0xFFFF << (10 - keep_explicit_bits)) & 0xFFFF
    m = m - 65536 if m >= 32768 else m
    return (t.view(torch.int16) & m).view(torch.float16)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=16384)
    ap.add_argument("--n", type=int, default=50000)
    ap.add_argument("--d", type=int, default=3072)
    ap.add_argument("--iters", type=int, default=60)
    ap.add_argument("--only", type=str, default="")   # comma list of variant indices
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
    prep = be.prepare_rows(x, a.m)
    smi = Smi()

    def run(qh, ql, yh, yl, precision="f16x3"):
        return be.posterior_stats(precision=precision, M=a.m, N=a.n, d=a.d, q_norm=prep["norms"], y_norm=y_norm,
                                  inv_temp=inv_t, q_split=(qh, ql, prep["inv_scale"]), y_split=(yh, yl),
                                  y_inv_scale=1.0 / scale, cta_group=2)

    variants = [
        ("real hi/lo (10+10 explicit bits)", 10, 10, "f16x3"),
        ("lo masked to 7 explicit bits", 10, 7, "f16x3"),
        ("lo masked to 4 explicit bits", 10, 4, "f16x3"),
        ("lo masked to 0 explicit bits (powers of two)", 10, 0, "f16x3"),
        ("hi masked to 7 bits (bf16-like), lo 7", 7, 7, "f16x3"),
        ("hi 7 bits, lo real", 7, 10, "f16x3"),
        ("real, single term f16x1", 10, 10, "f16x1"),
        ("real hi/lo again", 10, 10, "f16x3"),
        ("real, two terms f16x2", 10, 10, "f16x2"),
    ]
    if a.only:
        variants = [variants[int(i)] for i in a.only.split(",")]
    for name, hb, lb, prec in variants:
        qh, ql = mask16(prep["hi"], hb), mask16(prep["lo"], lb)
        yh, yl = mask16(ys["hi"], hb), mask16(ys["lo"], lb)
        run(qh, ql, yh, yl, prec)
        torch.cuda.synchronize()
        k = smi.mark()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters + 1)]
        evs[0].record()
        for i in range(a.iters):
            run(qh, ql, yh, yl, prec)
            evs[i + 1].record()
        torch.cuda.synchronize()
        each = [evs[i].elapsed_time(evs[i + 1]) for i in range(a.iters)]
        tail = each[len(each) // 2:]
        ms = sum(tail) / len(tail)
        terms = {"f16x3": 3, "f16x2": 2, "f16x1": 1}[prec]
        clk, pw = smi.since(k)
        print(f"{name:48s} {ms:8.3f} ms  executed {terms * 2 * a.d * a.m * a.n / ms / 1e9:7.1f} TFLOP/s  "
              f"sm {clk:6.0f} MHz  {pw:6.0f} W", flush=True)
    # zero operands: the power floor of the instruction stream itself
    z = torch.zeros_like(prep["hi"])
    zy = torch.zeros_like(ys["hi"])
    k = smi.mark()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters + 1)]
    evs[0].record()
    for i in range(a.iters):
        run(z, z, zy, zy)
        evs[i + 1].record()
    torch.cuda.synchronize()
    each = [evs[i].elapsed_time(evs[i + 1]) for i in range(a.iters)]
    tail = each[len(each) // 2:]
    ms = sum(tail) / len(tail)
    clk, pw = smi.since(k)
    print(f"{'all-zero operands':48s} {ms:8.3f} ms  executed {3 * 2 * a.d * a.m * a.n / ms / 1e9:7.1f} TFLOP/s  "
          f"sm {clk:6.0f} MHz  {pw:6.0f} W", flush=True)
    smi.proc.terminate()


if __name__ == "__main__":
    main()
