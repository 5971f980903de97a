# round 2, call 26: where the HOST time of a production-size drop-in run goes - 10 ms SIGPROF sampling of the arenes28_M400 sweeps
# (28 orbitals, M = 200 -> 400), aggregated for the library (planner) and for the reference binary (its own host code + the hooks)
O=gpurun_out/r2_26; mkdir -p $O
gcc -O2 -shared -fPIC -o /tmp/sprof.so scripts/probes/sprof.c
python - <<'PY'
import os, subprocess, time, sys
sys.path.insert(0, "tests")
import numpy as np
z = np.load("tests/golden/dropin_cases.npz")
name = "arenes28_M400"
work = "/tmp/prof_case"; os.makedirs(work, exist_ok=True)
for f in z[name + "/files"]:
    open(os.path.join(work, str(f)), "wb").write(z["%s/file/%s" % (name, f)].tobytes())
open(os.path.join(work, "dmrg.conf"), "wb").write(z[name + "/conf"].tobytes())
env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1", LD_PRELOAD="/tmp/sprof.so", SPROF_OUT="/tmp/sprof_dropin.out", B2D_DROPIN_STATS="/tmp/prof_case/stats.txt", B2D_DROPIN_TIMING="1")
t0 = time.time()
r = subprocess.run([os.path.abspath("oracle/_ref/block_gpu"), "dmrg.conf"], cwd=work, env=env, capture_output=True, text=True)
print("exit", r.returncode, "wall %.1f s" % (time.time() - t0))
print("\n".join(l for l in r.stderr.splitlines() if "B2D_TIMING" in l))
PY
python scripts/probes/sprof_aggregate.py /tmp/sprof_dropin.out $PWD/block_b200/lib/libblockb200.so 2>/dev/null | cut -c1-170 | head -60 > $O/host_profile_library.txt
python scripts/probes/sprof_aggregate.py /tmp/sprof_dropin.out $PWD/oracle/_ref/block_gpu 2>/dev/null | cut -c1-170 | head -80 > $O/host_profile_binary.txt
head -30 $O/host_profile_library.txt; grep -A25 "self top" $O/host_profile_binary.txt
