python -m pytest tests -m gpu -q 2>&1 | tail -2
for M in 172032 57344; do for RS in 0 1; do
  echo "== M=$M ROUND_SYNC=$RS"
  PDM_ROUND_SYNC=$RS python tools/quick_perf.py --m $M --iters 16 --configs 2:0:0 2>&1 | tail -1
  PDM_ROUND_SYNC=$RS ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:fused_gemm -s 1 -c 1 python tools/quick_perf.py --m $M --iters 2 --configs 2:0:0 2>&1 | grep -E "dram__bytes_read.sum|lts__t_sector_hit" | awk '{print "     ", $1, $2, $3}'
done; done
