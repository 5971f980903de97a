# Round 2, GPU call 1: is the device eigen-decomposition the outlier because of cuSOLVER's divide & conquer, or would the hand-written
# one-sided Jacobi at ALL sector sizes reproduce the reference's sweeps within the reference-vs-reference spread
# (tests/golden/eigvar_spread.npz)?   gpurun --timeout 1500 -- 'bash scripts/gpu_r2_eig_experiment.sh'
O=gpurun_out/r2_01
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tee $O/gpu.txt
for c in hubbard_L16_M1000 h2o_nosym_M60 h2o_nosym_M500; do
  timeout 900 python scripts/run_dropin_case.py $c B2D_DROPIN_OPTIONS=eig_jacobi_max=100000 --out $O/dropin 2>&1 | tee -a $O/jacobi_all.txt
done
for c in hubbard_L16_M1000 h2o_nosym_M60; do
  timeout 900 python scripts/run_dropin_case.py $c --out $O/dropin 2>&1 | tee -a $O/default.txt
done
