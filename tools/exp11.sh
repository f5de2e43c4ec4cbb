for rep in 1 2; do
for CFG in 2:16:4 2:14:5 2:12:6 2:9:8 2:8:9; do for HB in normal first; do
  echo "== rep $rep cfg $CFG HINT_B=$HB: $(PDM_HINT_B=$HB python tools/quick_perf.py --m 172032 --iters 16 --configs $CFG 2>&1 | tail -1 | cut -c1-90)"
done; done; done
