# Round 2, GPU call 8: why does the factorised drop-in fail on synthetic_16o_M300 (Davidson iteration cap)?
O=gpurun_out/r2_08
mkdir -p $O
for c in "synthetic_14o_M200 B2D_DROPIN_OPTIONS=factorised=1" "synthetic_14o_M200 B2D_DROPIN_OPTIONS=factorised=1,presum_identity=0" "synthetic_14o_M200 B2D_DROPIN_OPTIONS=factorised=1 B2D_DROPIN_CACHE=host" \
         "synthetic_14o_M200 B2D_DROPIN_OPTIONS=factorised=1 B2D_DROPIN_CHECK=1"; do
  timeout 600 python scripts/run_dropin_case.py $c --out $O/dropin 2>&1 | cut -c1-400 | head -12 | tee -a $O/dropin.txt
done
grep -h "B2D_CHECK" $O/dropin/*CHECK*.stderr.txt | head -40 | cut -c1-200 | tee $O/check.txt
timeout 600 python -m pytest tests/test_gpu_factorised.py tests/test_gpu_opbuild.py -m gpu -x -q 2>&1 | tail -6 | tee $O/pytest_new.txt
