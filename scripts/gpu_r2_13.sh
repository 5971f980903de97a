# Round 2, GPU call 13: token-only host copies in the cached drop-in - full GPU suite, all drop-in cases with per-sweep table, P4 stand-in at M <= 400 again
O=gpurun_out/r2_13
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $O/pytest_gpu.txt
for c in c2_d2h_M50 c2_d2h_M50_noise c2_d2h_M50_twodotnoise c2_d2h_M50_onedot_tail h2o_nosym_M60 h2o_nosym_M500 hubbard_L16_M80 hubbard_L16_M1000 synthetic_14o_M200 synthetic_16o_M300 arenes28_M400; do
  timeout 900 python scripts/run_dropin_case.py $c --out $O/dropin 2>&1 | tee -a $O/dropin_sweeps.txt | head -1
done
python - <<'PY'
import re, glob
for f in sorted(glob.glob("gpurun_out/r2_13/dropin/*.stats.txt")):
    tot = {}
    for l in open(f):
        for k, v in re.findall(r"(\w+)=([-\d.e+]+)", l): tot[k] = tot.get(k, 0) + float(v)
    print(f.split("/")[-1], {k: round(v, 2) for k, v in tot.items() if k.endswith("_s") or k in ("launches", "cache_uses")})
PY
B2D_EIG_DEBUG=1 timeout 600 python -m pytest tests/test_gpu_eig.py -m gpu -q -s 2>&1 | grep "b2d eig" | tee $O/eig_sweeps.txt
