#!/usr/bin/env python
"""Fixtures for the operator-construction oracle (oracle/opbuild_oracle.py): for chosen block iterations of the REAL
reference (oracle/_ref/block_dump with ORACLE_DUMP_CHILDREN=1) the two children of the enlarged left block with every operator they
carry ("LL." system, "LR." dot), the product StateInfo maps of the enlarged block ("L.si.*") and the enlarged block's operators as the
reference's own Op::build made them ("LA.").  Run in the build container only."""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import dumpio  # noqa: E402
import make_golden  # noqa: E402

CASES = {"c2_d2h_M30": [4, 13], "hubbard_M40": [7], "h2o_c1_M32": [9]}


def main():
    for name, calls in CASES.items():
        tdir, extra, conf, _ = make_golden.CASES[name]
        work = tempfile.mkdtemp(prefix="opbuild_")
        for f in ["FCIDUMP"] + extra:
            shutil.copy(os.path.join(make_golden.REF, "dmrg_tests", tdir, f), work)
        open(os.path.join(work, "dmrg.conf"), "w").write(conf)
        os.makedirs(os.path.join(work, "dump"))
        env = dict(os.environ, ORACLE_DUMP_DIR="dump", ORACLE_DUMP_CALLS=",".join(map(str, calls)), ORACLE_DUMP_CHILDREN="1",
                   OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1")
        out = subprocess.run([make_golden.DUMP, "dmrg.conf"], cwd=work, env=env, capture_output=True, text=True)
        if out.returncode != 0:
            print(out.stdout[-2000:], out.stderr[-2000:]); raise SystemExit("reference run failed for " + name)
        for c in calls:
            rec = dumpio.read_records(os.path.join(work, "dump", "site%d.bin" % c))
            keep = {k: v for k, v in rec.items() if k.startswith(("LL.", "LR.", "LA.", "L.si.")) or k in ("meta", "sym", "spin_orbs_symmetry", "L.q", "L.dims", "L.sites", "v1", "v2", "screen_tol")}
            dst = os.path.join(HERE, "opbuild_%s_call%d.npz" % (name, c))
            np.savez_compressed(dst, **keep)
            print(name, c, "LL ops %d, LR ops %d, LA ops %d, %.1f kB" % (int(rec["LL.nops"][0]), int(rec["LR.nops"][0]), int(rec["LA.nops"][0]), os.path.getsize(dst) / 1e3))
        shutil.rmtree(work)


if __name__ == "__main__":
    main()
