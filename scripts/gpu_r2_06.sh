# Round 2, GPU call 6: phase timing of the cached drop-in; new scatter kernel; factorised test; first full-size factorised bench (20|20)
O=gpurun_out/r2_06
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_factorised.py tests/test_gpu_opbuild.py tests/test_z_gpu_next_rows.py -m gpu -x -q 2>&1 | tail -6 | tee $O/pytest_new.txt
for c in "synthetic_14o_M200 B2D_DROPIN_TIMING=1" "synthetic_14o_M200 B2D_DROPIN_TIMING=1 B2D_DROPIN_CACHE=host"; do
  timeout 900 python scripts/run_dropin_case.py $c --out $O/dropin 2>&1 | grep -v "dE=[+-][0-9].[0-9]*e-1[0-9] " | tee -a $O/dropin.txt
done
grep -h B2D_TIMING $O/dropin/*.stderr.txt | tee $O/timing.txt
timeout 1500 python bench.py --no-sweep --no-block-iteration --steps 3 > $O/bench_factorised.json 2> $O/bench_factorised.err
tail -c 3000 $O/bench_factorised.json; tail -5 $O/bench_factorised.err
