python -m pytest tests -m gpu -x -q 2>&1 | tail -15
L=physics-of-diffusion-models_b200/lib
for lib in libpdm_stall libpdm_stall_d64; do for f in 1 2; do
echo "== lib=$lib FLUSH=$f"
PDM_B200_LIB=$L/$lib.so PDM_FLUSH_KB=$f python tools/stall_probe.py --iters 6
done; done
