# Round 2, GPU call 2: block Jacobi eigen-solver (tests + drop-in cases that were on cuSOLVER), defaults flipped (device guess, batched opbuild)
O=gpurun_out/r2_02
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_eig.py -m gpu -x -q -s 2>&1 | tail -15 | tee $O/pytest_eig.txt
timeout 900 python -m pytest tests/test_gpu_hotpath.py -m gpu -x -q 2>&1 | tail -5 | tee $O/pytest_hotpath.txt
for c in hubbard_L16_M1000 h2o_nosym_M500 synthetic_14o_M200 c2_d2h_M50_onedot_tail; do
  timeout 900 python scripts/run_dropin_case.py $c --out $O/dropin 2>&1 | tee -a $O/dropin.txt
done
python - <<'PY'
import re, glob
for f in sorted(glob.glob("gpurun_out/r2_02/dropin/*.stats.txt")):
    tot = {}
    for l in open(f):
        for k, v in re.findall(r"(\w+)=([-\d.e+]+)", l): tot[k] = tot.get(k, 0) + float(v)
    print(f.split("/")[-1], {k: round(v, 2) for k, v in tot.items() if k.endswith("_s") or k == "launches"})
PY
