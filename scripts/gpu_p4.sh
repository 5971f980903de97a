# P4 stand-in (BASELINE configs[3]: "C2/D2h full-valence ..., M=2000" - the file is not in the reference tree; SURVEY 8d names
# dmrg_tests/dmrg_parameters/arenes/28_28_fie as the stand-in): 28 orbitals / 28 electrons, C1, two-dot sweeps up to M = 2000 through the
# drop-in on ONE GPU.  The CPU reference cannot run this M here; its M <= 400 sweeps are the golden case arenes28_M400.
#   gpurun --timeout 2400 -- 'bash scripts/gpu_p4.sh'
#   P4_OPTIONS="factorised=1" P4_OUT=gpurun_out/p4_fact bash scripts/gpu_p4.sh      (library options through B2D_DROPIN_OPTIONS)
#   P4_GPUS=2 P4_OUT=gpurun_out/p4_n2 bash scripts/gpu_p4.sh                          (one process per GPU: the hooks read RANK / WORLD_SIZE /
#                                                                                      LOCAL_RANK, share the NCCL id through a file; every rank runs the same sweep)
O=${P4_OUT:-gpurun_out/p4}
N=${P4_GPUS:-1}
R=${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p $R/$O /tmp/p4 && cd /tmp/p4
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
rm -f /tmp/p4/nccl_id
T0=$(date +%s.%N)
PIDS=""
for r in $(seq 0 $((N - 1))); do
  mkdir -p /tmp/p4/r$r && cp /tmp/p4/FCIDUMP /tmp/p4/dmrg.conf /tmp/p4/r$r/
  ( cd /tmp/p4/r$r && env OPENBLAS_NUM_THREADS=1 OMP_NUM_THREADS=1 B2D_DROPIN_STATS=/tmp/p4/r$r/stats.txt B2D_DROPIN_TIMING=1 B2D_DROPIN_OPTIONS="${P4_OPTIONS:-}" \
      $( [ $N -gt 1 ] && echo "RANK=$r WORLD_SIZE=$N LOCAL_RANK=$r B2D_NCCL_ID_FILE=/tmp/p4/nccl_id" ) \
      stdbuf -oL $R/oracle/_ref/block_gpu dmrg.conf > $R/$O/stdout_r$r.txt 2> $R/$O/stderr_r$r.txt ) &
  PIDS="$PIDS $!"
done
RC=0; for p in $PIDS; do wait $p || RC=$?; done
T1=$(date +%s.%N)
cd $R
for r in $(seq 0 $((N - 1))); do echo "rank $r"; grep -E "Sweep Energy|Elapsed Sweep Wall" $O/stdout_r$r.txt; done | tee $O/sweeps.txt
echo "exit $RC, $N GPU(s), options '${P4_OPTIONS:-}', total wall $(python -c "print('%.1f' % ($T1 - $T0))") s" | tee -a $O/sweeps.txt
grep B2D_TIMING $O/stderr_r0.txt | tee -a $O/sweeps.txt
cp /tmp/p4/r0/stats.txt $O/stats.txt
for r in $(seq 0 $((N - 1))); do grep -E "Block Iteration|# states|Sweep Energy|Elapsed" $O/stdout_r$r.txt | cut -c1-200 > $O/stdout_short_r$r.txt; rm -f $O/stdout_r$r.txt; tail -c 8000 $O/stderr_r$r.txt > $O/stderr_tail_r$r.txt; rm -f $O/stderr_r$r.txt; done
P4_STATS=$O/stats.txt python - <<'PY'
import os, re
rows = [dict((k, float(v)) for k, v in re.findall(r"(\w+)=([-\d.e+]+)", l)) for l in open(os.environ["P4_STATS"])]
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
