# Final round-1 verification + the last profile: GPU tests, smoke, default bench, ncu of the E4M3 screening kernel.
set -x
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final.log
tail -3 gpurun_out/pytest_gpu_final.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_final.log
tail -2 gpurun_out/smoke_final.log
timeout 600 python bench.py > gpurun_out/bench_final3.json 2> gpurun_out/bench_final3.err; echo "bench rc=$?"
timeout 300 python tools/quick_screen.py --n 50000 --d 3072 --b 1024 --nt 170 --tmin 1e-4 --tmax 0.5 --iters 1 > gpurun_out/quick_f8.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_gemm -s 3 -c 1 \
    -o gpurun_out/prof_r1_fused_f8_block python tools/quick_screen.py --n 50000 --d 3072 --b 1024 --nt 170 --tmin 1e-4 --tmax 0.5 --iters 1 > gpurun_out/ncu_full_f8.log 2>&1
ncu -i gpurun_out/prof_r1_fused_f8_block.ncu-rep --page raw --csv 2>/dev/null | head -3 | cut -c1-400
python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_final3.json"))
print("headline", j["value"], j["ms_per_step"], j["e2e"]["value"], j["roofline"]["frac"], j["gpu_launches"])
s = j["screened"]; print("screened", s["value"], s["ms_per_step"], s.get("roofline_frac"))
s = j["lattice_8bit"]["screened"]; print("lattice", j["lattice_8bit"]["value"], "screened", s["value"], s["ms_per_step"], s.get("roofline_frac"))
print(j["denoiser_step"]); print(j["clocks"], j["cpu_baseline"])
PY
