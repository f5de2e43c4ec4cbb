L=physics-of-diffusion-models_b200/lib
for h in 20000 2000 200 0; do
echo "== WAIT_HINT=$h"
PDM_WAIT_HINT_NS=$h PDM_B200_LIB=$L/libpdm_stall.so python tools/stall_probe.py --iters 5 --precs f16x3,f16x2 2>&1 | grep -v "^max"
done
