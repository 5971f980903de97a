# P4 stand-in (BASELINE configs[3]: "C2/D2h full-valence ..., M=2000" - the file is not in the reference tree; SURVEY 8d names
# dmrg_tests/dmrg_parameters/arenes/28_28_fie as the stand-in): 28 orbitals / 28 electrons, C1, two-dot sweeps up to M = 2000 through the
# drop-in on ONE GPU.  The CPU reference cannot run this M here; its M <= 400 sweeps are the golden case arenes28_M400.
#   gpurun --timeout 2400 -- 'bash scripts/gpu_p4.sh'
O=gpurun_out/p4
mkdir -p $O /tmp/p4 && cd /tmp/p4
python - <<'PY'
import numpy as np, os
z = np.load(os.path.join(os.environ.get("GRAFT_REPO_ROOT", "/root/repo"), "tests/golden/dropin_cases.npz"))
open("FCIDUMP", "wb").write(z["arenes28_M400/file/FCIDUMP"].tobytes())
open("dmrg.conf", "w").write("""nelec 28
spin 0
irrep 1
hf_occ integral
schedule
0 200 1.0e-6 1.0e-4
2 400 1.0e-6 0.0
4 1000 1.0e-6 0.0
6 2000 1.0e-6 0.0
end
maxiter 8
twodot
sweep_tol 1e-12
orbitals FCIDUMP
noreorder
outputlevel 0
warmup local_2site
""")
PY
R=${GRAFT_REPO_ROOT:-/root/repo}
T0=$(date +%s.%N); env OPENBLAS_NUM_THREADS=1 OMP_NUM_THREADS=1 B2D_DROPIN_STATS=/tmp/p4/stats.txt B2D_DROPIN_TIMING=1 stdbuf -oL $R/oracle/_ref/block_gpu dmrg.conf > $R/$O/stdout.txt 2> $R/$O/stderr.txt
T1=$(date +%s.%N)
cd $R
grep -E "Sweep Energy|Elapsed Sweep Wall" $O/stdout.txt | tee $O/sweeps.txt
echo "total wall $(python -c "print('%.1f' % ($T1 - $T0))") s" | tee -a $O/sweeps.txt
grep B2D_TIMING $O/stderr.txt | tee -a $O/sweeps.txt
cp /tmp/p4/stats.txt $O/stats.txt
python - <<'PY'
import re
rows = [dict((k, float(v)) for k, v in re.findall(r"(\w+)=([-\d.e+]+)", l)) for l in open("gpurun_out/p4/stats.txt")]
tot = {}
for r in rows:
    for k, v in r.items(): tot[k] = tot.get(k, 0) + v
print({k: round(v, 2) for k, v in tot.items() if k.endswith("_s") or k in ("launches", "cache_uses", "n_multiply")})
big = sorted(rows, key=lambda r: -r.get("sigma_flops", 0))[:5]
for r in big:
    gf = r["sigma_flops"] * r["n_multiply"] / max(r["davidson_dev_ms"], 1e-9) / 1e6
    print("heaviest block iterations: lsites %d W %d sigma_flops %.3e n_multiply %d davidson_dev_ms %.1f -> %.0f GFLOP/s (algorithmic, whole Davidson solve)" %
          (r["lsites"], r["W"], r["sigma_flops"], r["n_multiply"], r["davidson_dev_ms"], gf))
PY
