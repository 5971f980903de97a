mkdir -p gpurun_out/r6
for c in "h2o_nosym_M60 B2D_DROPIN_EIG=host" "h2o_nosym_M500 B2D_DROPIN_EIG=host" "hubbard_L16_M1000 B2D_DROPIN_EIG=host" "h2o_nosym_M60 B2D_DROPIN_OPTIONS=eig_jacobi_max=4096" "hubbard_L16_M80" ; do timeout 900 python scripts/run_dropin_case.py $c --out gpurun_out/r6/dropin 2>&1 | tee -a gpurun_out/r6/dropin_summary.txt; done
python bench.py --no-cpu --no-block-iteration --sweep-case synthetic_18o_M500 --steps 2 --warmup 3 > gpurun_out/r6/bench_sweep18.json 2> gpurun_out/r6/bench_sweep18.err; python -c "
import json; l=json.loads(open('gpurun_out/r6/bench_sweep18.json').read().strip().splitlines()[-1]); print(json.dumps(l['sweep'])); print(l['value'], l['roofline']['frac'])"; tail -3 gpurun_out/r6/bench_sweep18.err
