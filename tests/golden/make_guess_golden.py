#!/usr/bin/env python
"""Fixtures for the guess-wavefunction transform (GuessWave::transform_previous_wavefunction,
guess_wavefunction.C:524-636, two-dot branch): for chosen block iterations of the REAL reference (oracle/_ref/block_dump with
ORACLE_DUMP_GUESS=1) and every root, the previous wavefunction and its StateInfo tree as the reference loads them from its scratch
files, the two rotation matrices, every StateInfo table the transform reads, and the reference's own trial vector ("gw<i>.trial").
Run in the build container only."""
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

# calls chosen after a run with ORACLE_DUMP_CALLS=all: one forward and one backward block iteration with a TRANSFORM guess
CASES = {"c2_d2h_M30": None, "hubbard_M40": None, "h2o_c1_M32": None}
FIRST = {"c2_d2h_M30": 9, "hubbard_M40": 6, "h2o_c1_M32": 10}   # skip the warm-up sweep (BASIC guesses)


def main():
    for name in CASES:
        tdir, extra, conf, _ = make_golden.CASES[name]
        work = tempfile.mkdtemp(prefix="guess_")
        for f in ["FCIDUMP"] + extra:
            shutil.copy(os.path.join(make_golden.REF, "dmrg_tests", tdir, f), work)
        open(os.path.join(work, "dmrg.conf"), "w").write(conf)
        os.makedirs(os.path.join(work, "dump"))
        env = dict(os.environ, ORACLE_DUMP_DIR="dump", ORACLE_DUMP_CALLS=",".join(map(str, range(FIRST[name], 60))), ORACLE_DUMP_GUESS="1", OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1")
        out = subprocess.run([make_golden.DUMP, "dmrg.conf"], cwd=work, env=env, capture_output=True, text=True)
        if out.returncode != 0:
            print(out.stdout[-2000:], out.stderr[-2000:]); raise SystemExit("reference run failed for " + name)
        have = []
        for c in range(60):
            path = os.path.join(work, "dump", "site%d.bin" % c)
            if not os.path.exists(path):
                continue
            rec = dumpio.read_records(path)
            if "gw.nroots" in rec:
                have.append((c, int(rec["meta"][1]), rec))     # meta[1]: forward sweep (left block starts at site 0)
        fwd = [h for h in have if h[1]]
        bwd = [h for h in have if not h[1]]
        print(name, "calls with a TRANSFORM guess:", [h[0] for h in have])
        picks = []
        if fwd: picks.append(fwd[len(fwd) // 2])
        if bwd: picks.append(bwd[len(bwd) // 2])
        for c, forward, rec in picks:
            keep = {k: v for k, v in rec.items() if k.startswith("gw") or k in ("meta", "psi_dq")}
            dst = os.path.join(HERE, "guess_%s_call%d.npz" % (name, c))
            np.savez_compressed(dst, **keep)
            print("  ", name, c, "forward" if forward else "backward", "roots %d, W = %d, %.1f kB" % (int(rec["gw.nroots"][0]), rec["gw0.trial"].size, os.path.getsize(dst) / 1e3))
        shutil.rmtree(work)


if __name__ == "__main__":
    main()
