for H in "normal normal" "last normal" "last first" "normal first"; do set -- $H
  echo "== HINT_A=$1 HINT_B=$2"
  PDM_HINT_A=$1 PDM_HINT_B=$2 python tools/quick_perf.py --m 172032 --iters 16 --configs 2:0:0 2>&1 | tail -1
  PDM_HINT_A=$1 PDM_HINT_B=$2 ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:fused_gemm -s 1 -c 1 python tools/quick_perf.py --m 172032 --iters 2 --configs 2:0:0 2>&1 | grep -E "dram__bytes_read.sum|lts__t_sector_hit" | awk '{print "     ", $1, $2, $3}'
done
for CFG in 2:18:4 2:14:5 2:12:6; do
  echo "== cfg $CFG"
  python tools/quick_perf.py --m 172032 --iters 16 --configs $CFG 2>&1 | tail -1
done
