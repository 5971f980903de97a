#!/usr/bin/env python
"""Fixtures for the ONE-DOT branch of the guess-wavefunction transform (guess_wavefunction.C:600-628 ->
GuessWave::onedot_transform_wavefunction :832-936): block iterations of the REAL reference run with the `onedot` keyword
(oracle/_ref/block_dump, ORACLE_DUMP_GUESS=1), one with the dot on the system side (transpose_guess_wave: rotate, then shuffle the dot
from the environment to the system) and one with the dot on the environment side (rotate only).  Run in the build container only."""
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

CASES = ["c2_d2h_M30", "h2o_c1_M32"]


def main():
    for name in CASES:
        tdir, extra, conf, _ = make_golden.CASES[name]
        conf = conf.replace("twodot\n", "onedot\n")
        work = tempfile.mkdtemp(prefix="guess1_")
        for f in ["FCIDUMP"] + extra:
            shutil.copy(os.path.join(make_golden.REF, "dmrg_tests", tdir, f), work)
        open(os.path.join(work, "dmrg.conf"), "w").write(conf)
        os.makedirs(os.path.join(work, "dump"))
        env = dict(os.environ, ORACLE_DUMP_DIR="dump", ORACLE_DUMP_CALLS=",".join(map(str, range(10, 80))), ORACLE_DUMP_GUESS="1", ORACLE_DUMP_ONLY_GUESS="1",
                   OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1")
        out = subprocess.run([make_golden.DUMP, "dmrg.conf"], cwd=work, env=env, capture_output=True, text=True)
        if out.returncode != 0:
            print(out.stdout[-2000:], out.stderr[-2000:]); raise SystemExit("reference run failed for " + name)
        have = {0: [], 1: []}
        for c in range(10, 80):
            path = os.path.join(work, "dump", "guess1dot_%d.bin" % c)
            if os.path.exists(path):
                rec = dumpio.read_records(path)
                have[int(rec["meta"][3])].append((c, rec))
        print(name, "one-dot TRANSFORM guesses: dot on the system side", [c for c, _ in have[1]], "on the environment side", [c for c, _ in have[0]])
        tr = []
        for c in range(10, 80):
            path = os.path.join(work, "dump", "guessT1_%d.bin" % c)
            if os.path.exists(path):
                tr.append((c, dumpio.read_records(path)))
        print(name, "one-dot TRANSPOSE guesses at calls", [c for c, _ in tr])
        for c, rec in tr[:2]:
            dst = os.path.join(HERE, "guessT1_%s_call%d.npz" % (name, c))
            np.savez_compressed(dst, **rec)
            print("  ", name, c, "one-dot transpose", "forward" if int(rec["meta"][1]) else "backward", "W = %d, %.1f kB" % (rec["gw0.trial"].size, os.path.getsize(dst) / 1e3))
        for flag in (0, 1):
            if not have[flag]:
                continue
            c, rec = max(have[flag], key=lambda h: h[1]["gw0.trial"].size)
            dst = os.path.join(HERE, "guess1dot_%s_call%d.npz" % (name, c))
            np.savez_compressed(dst, **rec)
            print("  ", name, c, "transpose_guess_wave =", flag, "roots %d, W = %d, %.1f kB" % (int(rec["gw.nroots"][0]), rec["gw0.trial"].size, os.path.getsize(dst) / 1e3))
        shutil.rmtree(work)


if __name__ == "__main__":
    main()
