"""CPU checks of bench.py's contract pieces that need no GPU: the reference arm's JSON line (keys the driver reads),
the DDPM temperature schedule against the oracle's restatement, and the parser of the committed ncu summary."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, PDM_BENCH_N="3000", PDM_BENCH_CPU_B="32", PDM_BENCH_CPU_NT="2", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "pairs/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and "workload" in line["config"]
    assert line["vs_baseline"] is None and line["gpu_launches"] == 0
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the arm must still use every host core, and report the stock
    # dataloader_batch_size = 100 variant beside the chunk-5000 one
    assert line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert line["cpu_baseline"]["stock_dataloader_batch_size_100"]["value"] > 0


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_schedule_and_profile_parser():
    sys.path.insert(0, ROOT)
    import bench
    from oracle import synthetic as syn
    t = bench.ddpm_temperatures(1000, 1e-4, 2.478e4)
    assert torch.equal(t, syn.ddpm_temperatures(1000))
    assert abs(t[0].item() - 1.0e-4) < 2e-5 and abs(t[-1].item() / 2.478e4 - 1) < 1e-5 and (t[1:] > t[:-1]).all()
    traffic = bench.profiled_traffic()
    assert traffic is None or 1e9 < traffic < 1e12          # bytes of one launch, from profiles/
    peaks = bench.measured_peaks()
    assert peaks["tflops"] > 100 and peaks["hbm_gbs"] > 1000
