#!/usr/bin/env python
"""Fixtures for the TRANSPOSE guess of the first block iteration of a sweep (GuessWave::transpose_previous_wavefunction,
guess_wavefunction.C:55-84, two-dot): the previous wavefunction as the REAL reference loads it and its own trial vector
(oracle/_ref/block_dump, ORACLE_DUMP_GUESS=1).  Run in the build container only."""
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


def main():
    for name in ("c2_d2h_M30", "h2o_c1_M32"):
        tdir, extra, conf, _ = make_golden.CASES[name]
        work = tempfile.mkdtemp(prefix="guessT_")
        for f in ["FCIDUMP"] + extra:
            shutil.copy(os.path.join(make_golden.REF, "dmrg_tests", tdir, f), work)
        open(os.path.join(work, "dmrg.conf"), "w").write(conf)
        os.makedirs(os.path.join(work, "dump"))
        env = dict(os.environ, ORACLE_DUMP_DIR="dump", ORACLE_DUMP_CALLS=",".join(map(str, range(8, 60))), ORACLE_DUMP_GUESS="1", OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1")
        out = subprocess.run([make_golden.DUMP, "dmrg.conf"], cwd=work, env=env, capture_output=True, text=True)
        if out.returncode != 0:
            print(out.stdout[-2000:], out.stderr[-2000:]); raise SystemExit("reference run failed for " + name)
        have = []
        for c in range(8, 60):
            path = os.path.join(work, "dump", "guessT_%d.bin" % c)
            if os.path.exists(path):
                have.append((c, dumpio.read_records(path)))
        print(name, "TRANSPOSE guesses at calls", [c for c, _ in have])
        for c, rec in have[:2]:
            dst = os.path.join(HERE, "guessT_%s_call%d.npz" % (name, c))
            np.savez_compressed(dst, **rec)
            print("  ", name, c, "forward" if int(rec["meta"][1]) else "backward", "roots %d, W = %d, %.1f kB" % (int(rec["gw.nroots"][0]), rec["gw0.trial"].size, os.path.getsize(dst) / 1e3))
        shutil.rmtree(work)


if __name__ == "__main__":
    main()
