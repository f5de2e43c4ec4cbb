# DRAM traffic per launch vs schedule plan at the 6 GiB block shape (172032 x 50000) and at 57344
for M in 172032 57344; do
for CFG in 2:17:4 2:16:4 2:14:5 2:12:6 2:10:7 2:8:9; do
  echo "== M=$M cfg=$CFG"
  python tools/quick_perf.py --m $M --iters 4 --configs $CFG 2>&1 | tail -1
  ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:fused_gemm -s 1 -c 1 python tools/quick_perf.py --m $M --iters 2 --configs $CFG 2>&1 | grep -E "dram__bytes_read.sum|lts__t_sector_hit" | awk '{print "     ", $1, $2, $3}'
done; done
