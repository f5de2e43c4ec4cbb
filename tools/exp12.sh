# f16x1 screening pass at the bench block shape: flush interval, schedule, L2 hints
set -x
for fk in 4 8 16; do
  PDM_FLUSH_KB=$fk python tools/quick_perf.py --m 172032 --precision f16x1 --iters 4 --configs "2:0:0,2:12:6,2:16:4,2:24:3,2:37:2,2:8:9" 2>&1 | grep -E "^cfg"
done
PDM_FLUSH_KB=8 PDM_HINT_B=normal python tools/quick_perf.py --m 172032 --precision f16x1 --iters 4 --configs "2:0:0" 2>&1 | grep -E "^cfg"
PDM_FLUSH_KB=8 PDM_SYNC_TILES=4 python tools/quick_perf.py --m 172032 --precision f16x1 --iters 4 --configs "2:0:0" 2>&1 | grep -E "^cfg"
PDM_FLUSH_KB=8 PDM_SYNC_TILES=16 python tools/quick_perf.py --m 172032 --precision f16x1 --iters 4 --configs "2:0:0" 2>&1 | grep -E "^cfg"
python tools/quick_perf.py --m 172032 --precision f16x3 --iters 3 --configs "2:0:0" 2>&1 | grep -E "^cfg"
