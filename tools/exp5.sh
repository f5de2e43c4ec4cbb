# L2 hints x schedule plans: steady-state time (no profiler) and DRAM bytes per launch (ncu metric pass)
run() { # name, env...
  echo "== $*"
  env "$@" python tools/quick_perf.py --m 57344 --iters 24 --configs $CFG 2>&1 | tail -1
  env "$@" ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:fused_gemm -s 1 -c 1 python tools/quick_perf.py --m 57344 --iters 2 --configs $CFG 2>&1 | grep -E "dram__bytes_read.sum|lts__t_sector_hit" | awk '{print "     ", $1, $2, $3}'
}
for CFG in 2:8:9 2:12:6 2:16:4 2:24:3; do
run PDM_HINT_A=normal PDM_HINT_B=normal
run PDM_HINT_A=last PDM_HINT_B=normal
run PDM_HINT_A=last PDM_HINT_B=first
run PDM_HINT_A=normal PDM_HINT_B=first
done
