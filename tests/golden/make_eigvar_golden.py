#!/usr/bin/env python
"""Reference-vs-reference spread of the per-sweep energies under a change of the density-matrix eigen-solver ONLY.

oracle/_ref/block_eigvar (oracle/eigvar_hooks.cpp) is the unmodified reference sweep with diagonalise_dm (rotationmat.C:258-279)
interposed at link time; ORACLE_EIGVAR picks dsyev (control), dsyevd, dsyevr (same OpenBLAS), ulp (dsyev on rho perturbed by one
rounding error per element) or jacobi.  Every drop-in case of tests/golden/dropin_cases.npz is run with each variant; the sweep
energies go to tests/golden/eigvar_spread.npz:

    <case>/variants        names
    <case>/energies        [variant, sweep-line] sweep energies
    <case>/spread          [sweep-line] max over variants of |E_variant - E_reference(dropin_cases.npz)|

tests/test_gpu_dropin.py bounds the GPU drop-in's per-sweep deviation by max(1e-8, this measured spread): where the reference itself
moves by more than 1e-8 Eh when only the basis inside (near-)degenerate eigenspaces of rho changes, no implementation whose rho or
eigen-solver differs in the last bit can be held to 1e-8 - and where it does not move, the GPU path is held to 1e-8.

Run in the build container only (needs `make -C oracle ref eigvar`)."""
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
BLOCK = os.path.join(ROOT, "oracle", "_ref", "block_eigvar")
CASES = os.path.join(HERE, "dropin_cases.npz")
OUT = os.environ.get("EIGVAR_OUT", os.path.join(HERE, "eigvar_spread.npz"))
SWEEP_RE = re.compile(r"M = (\d+)\s+state = (\d+)\s+Largest Discarded Weight = (\S+)\s+Sweep Energy = (\S+)")
VARIANTS = ["dsyev", "dsyevd", "dsyevr", "ulp", "jacobi"]


def run(z, name, variant):
    work = tempfile.mkdtemp(prefix="eigvar_")
    for f in z[name + "/files"]:
        open(os.path.join(work, str(f)), "wb").write(z["%s/file/%s" % (name, f)].tobytes())
    threads = int(os.environ.get("EIGVAR_THREADS", "1"))   # cases whose golden sweeps were made with several host threads (make_dropin_golden.py `threads`)
    open(os.path.join(work, "dmrg.conf"), "wb").write(z[name + "/conf"].tobytes() + (b"threads_per_node %d\n" % threads if threads > 1 else b""))
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS=str(threads), ORACLE_EIGVAR=variant)
    t0 = time.time()
    out = subprocess.run([BLOCK, "dmrg.conf"], cwd=work, env=env, capture_output=True, text=True)
    dt = time.time() - t0
    shutil.rmtree(work)
    if out.returncode != 0:
        print(out.stdout[-2000:], out.stderr[-2000:])
        raise SystemExit("%s / %s failed" % (name, variant))
    return [float(m.group(4)) for m in SWEEP_RE.finditer(out.stdout)], dt


def main():
    only = sys.argv[1:]
    z = np.load(CASES)
    names = sorted({k.split("/")[0] for k in z.files if k.endswith("/sweeps")})
    store = dict(np.load(OUT)) if os.path.exists(OUT) else {}
    for name in names:
        if only and name not in only:
            continue
        ref = [float(m.group(4)) for m in SWEEP_RE.finditer(z[name + "/sweeps"].tobytes().decode())]
        rows = []
        for v in VARIANTS:
            e, dt = run(z, name, v)
            assert len(e) == len(ref), (name, v, len(e), len(ref))
            rows.append(e)
            d = np.abs(np.array(e) - np.array(ref))
            print("%-24s %-7s %6.1f s  max|dE| = %.2e  final |dE| = %.2e   per sweep: %s" %
                  (name, v, dt, d.max(), d[-1], " ".join("%.1e" % x for x in d)), flush=True)
        rows = np.array(rows)
        store[name + "/variants"] = np.array(VARIANTS)
        store[name + "/energies"] = rows
        store[name + "/spread"] = np.abs(rows - np.array(ref)[None, :]).max(axis=0)
        np.savez_compressed(OUT, **store)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
