"""Print the numbers of a bench.py JSON line that DESIGN.md section 0 quotes (one line per entry)."""
import json
import sys

for path in sys.argv[1:]:
    try:
        j = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as exc:                      # noqa: BLE001
        print(path, "no bench line:", exc)
        continue
    print(path, "n_gpus", j["n_gpus"], "value %.4g" % j["value"], "ms/step %.2f" % j["ms_per_step"],
          "e2e %.4g" % j["e2e"]["value"], "frac %.4f" % j["roofline"]["frac"], "launches", j["gpu_launches"],
          j["config"].get("sharding"), j.get("clocks"))
    for k in ("parity", "denoiser_step", "c5_trajectory", "c3_hypersphere", "c4_celeba64", "screened"):
        print(" ", k, json.dumps(j.get(k))[:1600])
    lat = j.get("lattice_8bit") or {}
    print("  lattice_8bit", lat.get("value"), lat.get("roofline_frac"), (lat.get("screened") or {}).get("value"),
          (lat.get("screened") or {}).get("roofline_frac"))
