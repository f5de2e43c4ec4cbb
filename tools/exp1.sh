for hint in 20000 0; do for wide in 0 1; do
echo "== WAIT_HINT=$hint DRAIN_WIDE=$wide"
PDM_WAIT_HINT_NS=$hint PDM_DRAIN_WIDE=$wide python tools/power_probe.py --only 0,6,8 --iters 40
done; done
echo "== FLUSH_KB=2 wide=1"; PDM_FLUSH_KB=2 PDM_DRAIN_WIDE=1 python tools/power_probe.py --only 0,6,8 --iters 40
