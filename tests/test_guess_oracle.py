"""CPU: the numpy restatement of the two-dot guess-wavefunction transform (oracle/guess_oracle.py) against the trial
vectors of the REAL reference (tests/golden/guess_*.npz, tests/golden/make_guess_golden.py): GuessWave::transform_previous_wavefunction
(guess_wavefunction.C:524-636) on forward and backward block iterations of C2/D2h (two roots), H2O/C1 and Hubbard."""
import glob
import os

import numpy as np
import pytest

from oracle import guess_oracle as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "guess_*.npz")))


def test_fixtures_present():
    assert len(FIXTURES) >= 6
    directions = set()
    for f in FIXTURES:
        with np.load(f) as z:
            directions.add(int(z["meta"][1]))
            assert int(z["gw.nroots"][0]) >= 1
    assert directions == {0, 1}   # forward and backward sweeps


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(f)[:-4] for f in FIXTURES])
def test_transform_previous_wavefunction_matches_reference(path):
    rec = dict(np.load(path))
    for root in range(int(rec["gw.nroots"][0])):
        ref = rec["gw%d.trial" % root]
        got = G.transform_previous_wavefunction(rec, root)
        assert got.shape == ref.shape
        assert np.linalg.norm(ref) > 0.5          # a transformed, nearly normalised wavefunction, not a zero vector
        err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        assert err < 1e-13, (path, root, err)


@pytest.mark.parametrize("path", FIXTURES[:2], ids=[os.path.basename(f)[:-4] for f in FIXTURES[:2]])
def test_trial_layout_is_the_big_block_wavefunction(path):
    """The allowed (left sector, right sector) pairs of the trial vector are exactly dq.allow(q_left, q_right): the flat order the
    Davidson solver consumes (Wavefunction::FlattenInto)."""
    rec = dict(np.load(path))
    left, right = G.stateinfo(rec, "gw0.left."), G.stateinfo(rec, "gw0.right.")
    dq = rec["gw0.dq"][:3]
    mask = np.array([[G.allow(dq, left["q"][i], right["q"][j]) for j in range(len(right["dims"]))] for i in range(len(left["dims"]))], dtype=np.int32)
    assert (mask == rec["gw0.trial.allowed"]).all()


ONEDOT = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "guess1dot_*.npz")))


@pytest.mark.parametrize("path", ONEDOT, ids=[os.path.basename(f)[:-4] for f in ONEDOT])
def test_onedot_transform_matches_reference(path):
    """One-dot branch (GuessWave::onedot_transform_wavefunction, guess_wavefunction.C:832-936): dot on the system side (rotate, then
    shuffle the dot from the environment to the system) and dot on the environment side (rotate only)."""
    rec = dict(np.load(path))
    for root in range(int(rec["gw.nroots"][0])):
        ref = rec["gw%d.trial" % root]
        got = G.transform_previous_wavefunction_onedot(rec, root)
        assert got.shape == ref.shape and np.linalg.norm(ref) > 0.5
        assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-13, path


def test_onedot_fixtures_cover_both_dot_positions():
    flags = set()
    for f in ONEDOT:
        with np.load(f) as z:
            flags.add(int(z["gw.nroots"][1]))
    assert flags == {0, 1}


TRANSPOSE = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "guessT_*.npz")))


@pytest.mark.parametrize("path", TRANSPOSE, ids=[os.path.basename(f)[:-4] for f in TRANSPOSE])
def test_transpose_guess_matches_reference(path):
    """GuessWave::transpose_previous_wavefunction (guess_wavefunction.C:55-84): bit-exact."""
    rec = dict(np.load(path))
    for root in range(int(rec["gw.nroots"][0])):
        assert np.array_equal(G.transpose_previous_wavefunction(rec, root), rec["gw%d.trial" % root])


TRANSPOSE1 = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "guessT1_*.npz")))


@pytest.mark.parametrize("path", TRANSPOSE1, ids=[os.path.basename(f)[:-4] for f in TRANSPOSE1])
def test_onedot_transpose_guess_matches_reference(path):
    """GuessWave::onedot_transpose_wavefunction (guess_wavefunction.C:140-198)."""
    rec = dict(np.load(path))
    assert TRANSPOSE1
    for root in range(int(rec["gw.nroots"][0])):
        ref = rec["gw%d.trial" % root]
        assert np.linalg.norm(G.onedot_transpose_wavefunction(rec, root) - ref) / np.linalg.norm(ref) < 1e-13
