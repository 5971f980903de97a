# Round 2, GPU call 3: block Jacobi with noise floor; factorised enlarged-block operators (opbuild tests + drop-in sweeps)
O=gpurun_out/r2_03
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_eig.py -m gpu -q 2>&1 | tail -15 | tee $O/pytest_eig.txt
timeout 900 python -m pytest tests/test_gpu_opbuild.py -m gpu -x -q 2>&1 | tail -15 | tee $O/pytest_opbuild.txt
for c in "c2_d2h_M50_noise B2D_DROPIN_CHECK=1 B2D_DROPIN_OPTIONS=factorised=1" "synthetic_14o_M200 B2D_DROPIN_OPTIONS=factorised=1" "hubbard_L16_M80 B2D_DROPIN_OPTIONS=factorised=1" \
         "c2_d2h_M50_onedot_tail B2D_DROPIN_OPTIONS=factorised=1" "h2o_nosym_M500" "h2o_nosym_M500 B2D_DROPIN_OPTIONS=factorised=1"; do
  timeout 900 python scripts/run_dropin_case.py $c --out $O/dropin 2>&1 | tee -a $O/dropin.txt
done
grep -h "B2D_CHECK" $O/dropin/c2_d2h_M50_noise*stderr.txt | awk '{print $3}' | sort | uniq -c | tee $O/check_kinds.txt
grep -h "B2D_CHECK" $O/dropin/c2_d2h_M50_noise*stderr.txt | grep -o "max_abs_diff=[^ ]*" | cut -d= -f2 | sort -g | tail -3 | tee -a $O/check_kinds.txt
python - <<'PY'
import re, glob
for f in sorted(glob.glob("gpurun_out/r2_03/dropin/*.stats.txt")):
    tot = {}
    for l in open(f):
        for k, v in re.findall(r"(\w+)=([-\d.e+]+)", l): tot[k] = tot.get(k, 0) + float(v)
    print(f.split("/")[-1], {k: round(v, 2) for k, v in tot.items() if k.endswith("_s") or k == "launches"})
PY
