mkdir -p gpurun_out/r4
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r4/pytest_gpu.txt
cat gpurun_out/r4/pytest_gpu.txt
for c in "hubbard_L16_M1000 B2D_DROPIN_CHECK=1" hubbard_L16_M1000 h2o_nosym_M500 synthetic_14o_M200; do timeout 900 python scripts/run_dropin_case.py $c --out gpurun_out/r4/dropin 2>&1 | tee -a gpurun_out/r4/dropin_summary.txt; done
for f in gpurun_out/r4/dropin/*CHECK*.stderr.txt; do echo $f; grep select_states $f | grep -v "count=0" | head; grep diagonalise_dm $f | sort -t= -k5 -g | tail -3; done
python bench.py > gpurun_out/r4/bench_n1.json 2> gpurun_out/r4/bench_n1.err; tail -c 2500 gpurun_out/r4/bench_n1.json; tail -3 gpurun_out/r4/bench_n1.err
