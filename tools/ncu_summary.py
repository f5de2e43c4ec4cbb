"""Turn ncu outputs (launch list CSV, .ncu-rep raw page) into the small text summaries kept under profiles/."""
import collections
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__cluster_size", "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor", "smsp__cycles_active.avg",
    "smsp__warp_issue_stalled", "sm__cycles_active.avg",
]


def launch_list(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    tot = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row.get("Metric Unit", "ns"), 1.0)
        t = tot[row.get("Kernel Name", "?")]
        t[0] += 1
        t[1] += v
    total = sum(v[1] for v in tot.values())
    out = [f"# launch list {path}: {sum(v[0] for v in tot.values())} launches, {total / 1e6:.3f} ms device time "
           "(ncu: cold cache, serialised -- compare shares)", "ms,share_pct,launches,kernel"]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{v[1] / 1e6:.3f},{100 * v[1] / total:.2f},{v[0]},{k[:140]}")
    return "\n".join(out)


def raw_page(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        name = dict(zip(hdr, r)).get("Kernel Name", "?")
        out.append(f"# {rep}: kernel {name[:120]}")
        for h, u, v in zip(hdr, units, r):
            if any(h == k or h.startswith(k) for k in KEEP):
                out.append(f"{h},{v},{u}")
    return "\n".join(out)


if __name__ == "__main__":
    for arg in sys.argv[1:]:
        print(launch_list(arg) if arg.endswith(".csv") else raw_page(arg))
