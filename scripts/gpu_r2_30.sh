# round 2, call 30: templates evaluate the coefficients in the reference's order again - the threshold-regime sweeps, operator construction
mkdir -p gpurun_out/r2_30
timeout 900 python -m pytest tests/test_gpu_dropin.py -m gpu -q -k "threshold_regime or (match_reference and (h2o_nosym_M500 or hubbard_L16_M1000 or h2o_nosym_M60 or c2_d2h_M50_twodotnoise))" 2>&1 | tail -6 | tee gpurun_out/r2_30/pytest.txt
timeout 900 python -m pytest tests/test_gpu_opbuild.py tests/test_gpu_factorised.py -m gpu -x -q 2>&1 | tail -3 | tee -a gpurun_out/r2_30/pytest.txt
