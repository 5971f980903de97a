# round 2, call 28: pinned staging of the scatter descriptors - parity of the operator construction and two sweeps, host profile again
mkdir -p gpurun_out/r2_28
timeout 900 python -m pytest tests/test_gpu_opbuild.py tests/test_z_gpu_next_rows.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2_28/pytest.txt
timeout 900 python -m pytest tests/test_gpu_dropin.py -m gpu -x -q -k "match_reference and (c2_d2h_M50_noise or hubbard_L16_M1000 or synthetic_14o)" 2>&1 | tail -3 | tee -a gpurun_out/r2_28/pytest.txt
sed -e 's#gpurun_out/r2_26#gpurun_out/r2_28#' scripts/gpu_r2_26.sh > /tmp/prof.sh; bash /tmp/prof.sh 2>&1 | head -24
