# Round 2, GPU call 14: scatter kernel (vectorised fast path) against the HBM roofline; block Jacobi convergence with the keep-threshold criterion
O=gpurun_out/r2_14
mkdir -p $O
B2D_EIG_DEBUG=1 timeout 600 python -m pytest tests/test_gpu_eig.py tests/test_gpu_opbuild.py tests/test_gpu_factorised.py tests/test_z_gpu_next_rows.py -m gpu -x -q -s 2>&1 | grep -E "b2d eig|passed|failed|Error" | tee $O/pytest.txt
timeout 300 python scripts/bench_next_rows.py 2>&1 | tail -1 > $O/bench_next_rows.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_14/bench_next_rows.json").read())
oc = d.get("operator_construction", {})
print(oc.get("roofline"), oc.get("batched"))
PY
for c in hubbard_L16_M1000 h2o_nosym_M500; do B2D_EIG_DEBUG=1 timeout 600 python scripts/run_dropin_case.py $c --out $O/dropin 2>&1 | grep -v "e-1[0-9] " | tee -a $O/dropin.txt; done
grep -h "b2d eig" $O/dropin/*.stderr.txt | sort | uniq -c | sort -rn | head -8
python - <<'PY'
import re, glob
for f in sorted(glob.glob("gpurun_out/r2_14/dropin/*.stats.txt")):
    tot = {}
    for l in open(f):
        for k, v in re.findall(r"(\w+)=([-\d.e+]+)", l): tot[k] = tot.get(k, 0) + float(v)
    print(f.split("/")[-1], {k: round(v, 2) for k, v in tot.items() if k.endswith("_s")})
PY
