# round 2, call 27: scatter-task / factor templates in the operator-construction planner - parity, then the host profile of call 26 again
mkdir -p gpurun_out/r2_27
timeout 900 python -m pytest tests/test_gpu_opbuild.py tests/test_gpu_factorised.py tests/test_z_gpu_next_rows.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2_27/pytest.txt
timeout 900 python -m pytest tests/test_gpu_dropin.py -m gpu -x -q -k "either_operator_form or (match_reference and (noise or hubbard_L16_M80 or h2o_nosym_M60))" 2>&1 | tail -4 | tee -a gpurun_out/r2_27/pytest.txt
sed -e 's#gpurun_out/r2_26#gpurun_out/r2_27#' scripts/gpu_r2_26.sh > /tmp/prof.sh; bash /tmp/prof.sh 2>&1 | head -40
