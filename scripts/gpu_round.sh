mkdir -p gpurun_out/r3
python -m pytest tests/test_gpu_dropin.py tests/test_gpu_hotpath.py -m gpu -q 2>&1 | tail -25 > gpurun_out/r3/pytest_gpu.txt
cat gpurun_out/r3/pytest_gpu.txt
for c in "h2o_nosym_M500 B2D_DROPIN_CHECK=1" "hubbard_L16_M1000 B2D_DROPIN_CHECK=1" c2_d2h_M50 hubbard_L16_M1000 synthetic_14o_M200 h2o_nosym_M500 "synthetic_14o_M200 B2D_DROPIN_TRANSFORM=reference"; do timeout 900 python scripts/run_dropin_case.py $c --out gpurun_out/r3/dropin 2>&1 | tee -a gpurun_out/r3/dropin_summary.txt; done
for f in gpurun_out/r3/dropin/*CHECK*.stderr.txt; do echo $f; grep select_states $f | grep -v "count=0" | head; grep diagonalise_dm $f | sort -t= -k5 -g | tail -3; done
