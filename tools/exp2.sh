L=physics-of-diffusion-models_b200/lib
for f in 1 2 4; do PDM_FLUSH_KB=$f python tools/accuracy_probe.py; done
for rep in 1 2; do
for lib in libpdm_b200 libpdm_d64; do for f in 1 2; do
echo "== lib=$lib FLUSH=$f rep=$rep"
PDM_B200_LIB=$L/$lib.so PDM_FLUSH_KB=$f python tools/power_probe.py --only 0,6,8 --iters 40
done; done; done
