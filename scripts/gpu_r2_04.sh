# Round 2, GPU call 4: Newton-Schulz clean-up of the block-Jacobi eigenvectors; device block cache (N3) default on; full GPU suite
O=gpurun_out/r2_04
mkdir -p $O
B2D_EIG_DEBUG=1 timeout 600 python -m pytest tests/test_gpu_eig.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -12 | tee $O/pytest_eig.txt
for c in c2_d2h_M50 c2_d2h_M50_noise c2_d2h_M50_onedot_tail hubbard_L16_M1000 synthetic_14o_M200 "synthetic_14o_M200 B2D_DROPIN_OPTIONS=factorised=1" "synthetic_14o_M200 B2D_DROPIN_CACHE=host" h2o_nosym_M60; do
  timeout 900 python scripts/run_dropin_case.py $c --out $O/dropin 2>&1 | grep -v "dE=[+-][0-9].[0-9]*e-1[0-9] " | tee -a $O/dropin.txt
done
python - <<'PY'
import re, glob
for f in sorted(glob.glob("gpurun_out/r2_04/dropin/*.stats.txt")):
    tot = {}
    for l in open(f):
        for k, v in re.findall(r"(\w+)=([-\d.e+]+)", l): tot[k] = tot.get(k, 0) + float(v)
    print(f.split("/")[-1], {k: round(v, 2) for k, v in tot.items() if k.endswith("_s") or k in ("launches", "cache_uses")})
PY
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $O/pytest_gpu.txt
