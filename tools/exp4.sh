L=physics-of-diffusion-models_b200/lib
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for rep in 1 2; do
for lib in libpdm_old libpdm_b200; do
echo "== quick_perf lib=$lib rep=$rep"
PDM_B200_LIB=$L/$lib.so python tools/quick_perf.py --m 57344 --iters 30 --configs 2:0:0 2>&1 | tail -2
done; done
for lib in libpdm_stall libpdm_stall_nospin; do for f in 1 2; do
echo "== lib=$lib FLUSH=$f"
PDM_B200_LIB=$L/$lib.so PDM_FLUSH_KB=$f python tools/stall_probe.py --iters 5 --precs f16x3,f16x2
done; done
