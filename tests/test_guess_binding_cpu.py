"""CPU: the reference-side binding of the guess-wavefunction transform (tests/dropin/guess_binding.hpp - the same marshalling the
drop-in binary uses) + the C++ planner (b2d_guess_plan) on EVERY guess of whole sweeps of the unmodified reference.

oracle/_ref/block_guesscheck (tests/dropin/guess_plan_cpu_check.cpp) is the reference sweep with one link-time wrap: after the
reference's own GuessWave::guess_wavefunctions has produced its trial vectors, the binding describes the same guess to a planning-only
context, the exported plan is executed with plain loops (the descriptors the device kernels read) and compared with the reference's
vector.  Covers all five forms (two-dot / one-dot TRANSFORM with either dot position, two-dot / one-dot TRANSPOSE), every root, every
block iteration, both sweep directions - far more sector structures than the golden fixtures."""
import collections
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "block_guesscheck")
CASES = os.path.join(ROOT, "tests", "golden", "dropin_cases.npz")


def run(name):
    z = np.load(CASES)
    work = tempfile.mkdtemp(prefix="guesscheck_" + name + "_")
    for f in z[name + "/files"]:
        open(os.path.join(work, str(f)), "wb").write(z["%s/file/%s" % (name, f)].tobytes())
    open(os.path.join(work, "dmrg.conf"), "wb").write(z[name + "/conf"].tobytes())
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1")
    out = subprocess.run([EXE, "dmrg.conf"], cwd=work, env=env, capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    by_mode, skipped, other = collections.defaultdict(list), [], []
    for l in out.stderr.splitlines():
        if not l.startswith("B2D_GUESSCHECK"):
            continue
        m = re.search(r"mode=(\d+) root=(\d+) W=(\d+) max_abs_diff=(\S+) max_abs=(\S+)", l)
        if m:
            by_mode[int(m.group(1))].append((float(m.group(4)), int(m.group(3)), float(m.group(5))))
        elif "skipped=" in l:
            skipped.append(l)
        else:
            other.append(l)
    return by_mode, skipped, other


@pytest.mark.parametrize("name,modes", [("hubbard_L16_M80", {0, 3}), ("c2_d2h_M50_onedot_tail", {0, 1, 2, 3, 4})])
def test_every_guess_of_a_sweep_is_reproduced(name, modes):
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/block_guesscheck not built (make -C oracle guesscheck; needs the reference sources)")
    by_mode, skipped, other = run(name)
    assert not other, other[:3]                 # no planner errors
    assert not skipped, skipped[:3]             # every TRANSFORM / TRANSPOSE guess of these runs is one of the covered forms
    assert set(by_mode) == modes, sorted(by_mode)
    worst = 0.0
    for mode, rows in by_mode.items():
        for diff, w, scale in rows:
            assert w > 0 and scale > 0
            assert diff <= 1e-13, (name, mode, diff)     # measured: <= 4.5e-16 (mode 3 is bit-exact)
            worst = max(worst, diff)
    print("%s: %s guesses, worst |difference| %.1e" % (name, {m: len(v) for m, v in sorted(by_mode.items())}, worst))
