# Round 2, GPU call 7: host-time profile of the drop-in sweep (10 ms SIGPROF sampler), full GPU suite, default bench run
O=gpurun_out/r2_07
mkdir -p $O
gcc -O2 -shared -fPIC -o /tmp/sprof.so scripts/probes/sprof.c
python - <<'PY'
import os, subprocess, sys, time
sys.path.insert(0, "tests")
import numpy as np
import test_gpu_dropin as T
z = np.load(T.CASES)
name = "synthetic_16o_M300"
work = "/tmp/prof_case"
os.makedirs(work, exist_ok=True)
for f in z[name + "/files"]:
    open(os.path.join(work, str(f)), "wb").write(z["%s/file/%s" % (name, f)].tobytes())
threads = os.cpu_count()
open(os.path.join(work, "dmrg.conf"), "w").write(z[name + "/conf"].tobytes().decode() + "threads_per_node %d\n" % threads)
env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS=str(threads), LD_PRELOAD="/tmp/sprof.so", SPROF_OUT="/tmp/sprof_dropin.out",
           B2D_DROPIN_STATS=os.path.join(work, "stats.txt"), B2D_DROPIN_TIMING="1")
t0 = time.time()
r = subprocess.run([T.BLOCK_GPU, "dmrg.conf"], cwd=work, env=env, capture_output=True, text=True)
print("drop-in under the sampler: exit %d, %.1f s wall, %d threads" % (r.returncode, time.time() - t0, threads))
print("\n".join(l for l in r.stdout.splitlines() if "Sweep Energy" in l))
print("\n".join(l for l in r.stderr.splitlines() if "B2D_TIMING" in l))
PY
python scripts/probes/sprof_aggregate.py /tmp/sprof_dropin.out $PWD/oracle/_ref/block_gpu 2>/dev/null | cut -c1-160 | head -70 | tee $O/dropin_host_profile.txt
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $O/pytest_gpu.txt
timeout 1500 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
tail -c 1500 $O/bench_n1.json; tail -3 $O/bench_n1.err
