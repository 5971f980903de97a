#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the REAL reference (oracle/_ref/block_dump = unmodified sanshar/Block objects +
oracle/ref_dump.cpp link-time hooks).  Run in the build container only (needs /root/reference and `make -C oracle ref`);
the GPU box uses the committed .npz files.

Each fixture is one SpinBlock::RenormaliseFrom call (renormalise.C:39) of a two-dot sweep: the two halves of the big
block with every hot-path operator, a deterministic psi and the reference's H.psi, diag(H), the Davidson guesses /
solutions / eigenvalues / H-application count, the reduced density matrix, the rotation matrices, the discarded weight
and the renormalised operators.
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import dumpio  # noqa: E402

REF = os.environ.get("BLOCK_REFERENCE", "/root/reference")
DUMP = os.path.join(ROOT, "oracle", "_ref", "block_dump")

CASES = {
    # name: (test dir, extra files, conf, calls to keep)
    "c2_d2h_M30": ("c2_d2h_smallM", ["reorder.dat"], """nelec 8
spin 0
irrep 1
hf_occ integral
schedule
0 30 1.0e-6 0.0
end
maxiter 3
twodot
sweep_tol 1e-9
sym d2h
orbitals FCIDUMP
nroots 2
weights 0.5 0.5
reorder reorder.dat
outputlevel 0
""", [4, 13]),
    "hubbard_M40": ("hubbard", [], """nelec 8
spin 0
hf_occ integral
schedule
0 40 1.0e-7 0.0
end
maxiter 3
twodot
sweep_tol 1e-9
orbitals FCIDUMP
noreorder
outputlevel 0
warmup local_2site
""", [7, 12]),
    # schedule noise > 0: pins DensityMatrix::add_onedot_noise (density.C:332-399)
    "c2_d2h_M30_noise": ("c2_d2h_smallM", ["reorder.dat"], """nelec 8
spin 0
irrep 1
hf_occ integral
schedule
0 30 1.0e-6 1.0e-3
end
maxiter 3
twodot
sweep_tol 1e-9
sym d2h
orbitals FCIDUMP
nroots 2
weights 0.5 0.5
reorder reorder.dat
outputlevel 0
""", [5, 12]),
    "h2o_c1_M32_noise": ("h2o_nosym", [], """nelec 10
spin 0
irrep 1
hf_occ integral
schedule
0 32 1.0e-7 1.0e-4
end
maxiter 3
twodot
sweep_tol 1e-9
orbitals FCIDUMP
noreorder
outputlevel 0
""", [14]),
    "h2o_c1_M32": ("h2o_nosym", [], """nelec 10
spin 0
irrep 1
hf_occ integral
schedule
0 32 1.0e-7 0.0
end
maxiter 3
twodot
sweep_tol 1e-9
orbitals FCIDUMP
noreorder
outputlevel 0
""", [15]),
}


def main():
    only = sys.argv[1:]
    for name, (tdir, extra, conf, calls) in CASES.items():
        if only and name not in only:
            continue
        work = tempfile.mkdtemp(prefix="golden_")
        for f in ["FCIDUMP"] + extra:
            shutil.copy(os.path.join(REF, "dmrg_tests", tdir, f), work)
        open(os.path.join(work, "dmrg.conf"), "w").write(conf)
        os.makedirs(os.path.join(work, "dump"))
        env = dict(os.environ, ORACLE_DUMP_DIR="dump", ORACLE_DUMP_CALLS=",".join(map(str, calls)), OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1")
        out = subprocess.run([DUMP, "dmrg.conf"], cwd=work, env=env, capture_output=True, text=True)
        if out.returncode != 0:
            print(out.stdout[-2000:], out.stderr[-2000:]); raise SystemExit("reference run failed for " + name)
        sweeps = [l.strip() for l in out.stdout.splitlines() if "Sweep Energy" in l]
        for c in calls:
            rec = dumpio.read_records(os.path.join(work, "dump", "site%d.bin" % c))
            dst = os.path.join(HERE, "%s_site%d.npz" % (name, c))
            np.savez_compressed(dst, **rec)
            print(name, c, "W=%d" % rec["rpsi"].size, "%.1f kB" % (os.path.getsize(dst) / 1e3))
        open(os.path.join(HERE, name + ".sweeps.txt"), "w").write("\n".join(sweeps) + "\n")
        shutil.rmtree(work)


if __name__ == "__main__":
    main()
