"""CPU: the C++ planner of the guess-wavefunction transform (block_b200/csrc/guess.hpp behind b2d_guess_plan) pinned
against the REAL reference.  The plan is what the device executes - grouped-GEMM segment / group descriptors for stages 1 and 3, the
scatter tasks of the shuffle with their parity x 6j coefficients, the padded layouts - so it is exported through the C ABI
(b2d_guess_plan_export) and executed here descriptor by descriptor with numpy on a planning-only context (no GPU, no oracle): the
result must be the reference's own trial vector (tests/golden/guess_*.npz)."""
import glob
import os

import numpy as np
import pytest

from block_b200 import hotpath

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "guess_*.npz")))

GSEG = np.dtype([("a", "<i8"), ("b", "<i8"), ("alpha", "<f8"), ("lda", "<i4"), ("ldb", "<i4"), ("k", "<i4"), ("a_base", "u1"), ("b_base", "u1"),
                 ("a_trans", "u1"), ("b_kmajor", "u1")], align=True)
GGROUP = np.dtype([("c", "<i8"), ("ldc", "<i4"), ("m", "<i4"), ("n", "<i4"), ("seg_begin", "<i4"), ("seg_end", "<i4"), ("kiters", "<i4"),
                   ("c_base", "u1"), ("accumulate", "u1"), ("pad0", "u1"), ("pad1", "u1")], align=True)
KRON = np.dtype([("a", "<i8"), ("b", "<i8"), ("dst", "<i8"), ("coef", "<f8")] + [(n, "<i4") for n in
                ("a_rows", "a_cols", "lda", "a_t", "b_rows", "b_cols", "ldb", "b_t", "row0", "col0", "ldd", "pad")], align=True)
BLOCK = np.dtype([("ref_off", "<i8"), ("dev_off", "<i8"), ("rows", "<i4"), ("cols", "<i4"), ("ld", "<i4"), ("pad", "<i4")], align=True)
BASE_WORK, BASE_DST, BASE_AUX = 2, 3, 4    # gemm_desc.h


def tables(rec, root):
    p = "gw%d." % root
    out = {}
    for name in hotpath.GuessTransform.NAMES:
        out[name] = {k[len(p + name + "."):]: rec[k] for k in rec if k.startswith(p + name + ".")}
    return out


def make(rec, root, device):
    p = "gw%d." % root
    return hotpath.GuessTransform(rec[p + "dq"][:3], tables(rec, root), rec[p + "old.allowed"], rec[p + "lrot.shape"][:, 1], rec[p + "rrot.shape"][:, 1],
                                  device=device)


def run_groups(segs, groups, bases):
    for g in groups:
        m, n = int(g["m"]), int(g["n"])
        acc = np.zeros((m, n))
        for s in segs[int(g["seg_begin"]):int(g["seg_end"])]:
            k, lda, ldb = int(s["k"]), int(s["lda"]), int(s["ldb"])
            A, B = bases[int(s["a_base"])], bases[int(s["b_base"])]
            a0, b0 = int(s["a"]), int(s["b"])
            opa = A[a0:a0 + k * lda].reshape(k, lda)[:, :m].T if s["a_trans"] else A[a0:a0 + m * lda].reshape(m, lda)[:, :k]
            opb = B[b0:b0 + n * ldb].reshape(n, ldb)[:, :k].T if s["b_kmajor"] else B[b0:b0 + k * ldb].reshape(k, ldb)[:, :n]
            acc += float(s["alpha"]) * (opa @ opb)
        C = bases[int(g["c_base"])]
        c0, ldc = int(g["c"]), int(g["ldc"])
        view = C[c0:c0 + m * ldc].reshape(m, ldc)[:, :n]
        view[...] = view + acc if g["accumulate"] else acc


def execute_plan(gt, old, lrot, rrot):
    sizes = gt.export(9).view("<i4")
    assert tuple(sizes) == (GSEG.itemsize, GGROUP.itemsize, KRON.itemsize, BLOCK.itemsize)
    head = gt.export(7)
    n_old, n_l, n_r = (int(x) for x in head[:12].view("<i4"))
    image_size, t1_size, work_size = (int(x) for x in head[16:40].view("<i8"))
    blocks = gt.export(6).view(BLOCK)
    assert len(blocks) == n_old + n_l + n_r
    image, work = np.zeros(max(image_size, 1)), np.zeros(max(work_size, 1))
    for k, b in enumerate(blocks):
        src = old if k < n_old else (lrot if k < n_old + n_l else rrot)
        r, c, ld = int(b["rows"]), int(b["cols"]), int(b["ld"])
        assert ld % 2 == 0 and ld >= c and int(b["dev_off"]) % 16 == 0        # 16-byte rows, 128-byte blocks
        image[int(b["dev_off"]):int(b["dev_off"]) + r * ld].reshape(r, ld)[:, :c] = src[int(b["ref_off"]):int(b["ref_off"]) + r * c].reshape(r, c)
    tb = gt.export(8).view(BLOCK)
    Wp = max(int(b["dev_off"]) + int(b["rows"]) * int(b["ld"]) for b in tb)
    dst = np.zeros(Wp)
    bases = {BASE_WORK: work, BASE_DST: dst, BASE_AUX: image}
    run_groups(gt.export(0).view(GSEG), gt.export(1).view(GGROUP), bases)
    run_groups(gt.export(10).view(GSEG), gt.export(11).view(GGROUP), bases)      # second batch (one-dot): reads what the first wrote
    tasks, per_round = gt.export(2).view(KRON), gt.export(3).view("<i4")
    assert per_round.sum() == len(tasks) == gt.shuffle_tasks and len(per_round) == gt.shuffle_rounds
    first = 0
    for n in per_round:
        written = set()
        for t in tasks[first:first + int(n)]:
            assert t["b"] == 0 and t["b_rows"] == 1 and t["b_cols"] == 1
            target = dst if int(t["pad"]) & 1 else work      # pad bit 0: the task writes the trial vector
            source = image if int(t["pad"]) & 2 else work    # pad bit 1: the task reads the input image (transpose guess)
            if not int(t["pad"]) & 2:
                assert t["a_t"] == 0 and int(t["a"]) + (int(t["a_rows"]) - 1) * int(t["lda"]) + int(t["a_cols"]) <= t1_size   # reads stage-1 output only
            assert int(t["pad"]) & 1 or int(t["dst"]) >= t1_size
            key = (int(t["dst"]), int(t["row0"]))
            assert key not in written, "two tasks of one round write the same destination rows"
            written.add(key)
            rows, cols, lda = int(t["a_rows"]), int(t["a_cols"]), int(t["lda"])
            a0 = int(t["a"])
            blk = source[a0:a0 + cols * lda].reshape(cols, lda)[:, :rows].T if t["a_t"] else np.stack([source[a0 + i * lda:a0 + i * lda + cols] for i in range(rows)])
            for i in range(rows):
                d0 = int(t["dst"]) + (int(t["row0"]) + i) * int(t["ldd"]) + int(t["col0"])
                target[d0:d0 + cols] += float(t["coef"]) * blk[i]
        first += int(n)
    run_groups(gt.export(4).view(GSEG), gt.export(5).view(GGROUP), bases)
    flat = np.zeros(gt.trial_size)
    for b in tb:
        r, c, ld = int(b["rows"]), int(b["cols"]), int(b["ld"])
        flat[int(b["ref_off"]):int(b["ref_off"]) + r * c] = dst[int(b["dev_off"]):int(b["dev_off"]) + r * ld].reshape(r, ld)[:, :c].ravel()
    return flat


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(f)[:-4] for f in FIXTURES])
def test_planned_transform_reproduces_the_reference_trial_vector(path):
    rec = dict(np.load(path))
    for root in range(int(rec["gw.nroots"][0])):
        p = "gw%d." % root
        gt = make(rec, root, device=-1)
        try:
            assert gt.trial_size == rec[p + "trial"].size
            assert gt.old_size == rec[p + "old.data"].size and gt.lrot_size == rec[p + "lrot.data"].size and gt.rrot_size == rec[p + "rrot.data"].size
            assert gt.shuffle_rounds <= 2          # a spin-1/2 dot recouples at most two intermediate spins into one sector
            got = execute_plan(gt, rec[p + "old.data"], rec[p + "lrot.data"], rec[p + "rrot.data"])
            ref = rec[p + "trial"]
            err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
            assert err < 1e-13, (path, root, err)
        finally:
            gt.close()


def test_no_cpu_fallback_and_argument_checks():
    rec = dict(np.load(FIXTURES[0]))
    gt = make(rec, 0, device=-1)
    try:
        with pytest.raises(hotpath.B2DError, match="no CUDA device"):
            gt.transform(rec["gw0.old.data"], rec["gw0.lrot.data"], rec["gw0.rrot.data"])
    finally:
        gt.close()
    bad = tables(rec, 0)
    bad["dot"] = dict(bad["dot"], dims=np.array([1, 2, 1], np.int32))
    with pytest.raises(hotpath.B2DError, match="dot sectors"):
        hotpath.GuessTransform(rec["gw0.dq"][:3], bad, rec["gw0.old.allowed"], rec["gw0.lrot.shape"][:, 1], rec["gw0.rrot.shape"][:, 1], device=-1)


@pytest.mark.parametrize("M", [60, 600])
def test_synthetic_case_plans_and_executes(M):
    """CPU: the synthetic sector tables bench.py's guess-transform leg uses are consistent (b2d_guess_plan accepts them) and the plan
    is linear and non-trivial when executed with numpy."""
    from block_b200 import synthetic
    dq, tables, allowed, lcols, rcols, old, lrot, rrot = synthetic.make_guess_case(16, 16, M, 8)
    gt = hotpath.GuessTransform(dq, tables, allowed, lcols, rcols, device=-1)
    try:
        assert (gt.old_size, gt.lrot_size, gt.rrot_size) == (old.size, lrot.size, rrot.size)
        a = execute_plan(gt, old, lrot, rrot)
        b = execute_plan(gt, 2.0 * old, lrot, rrot)
        assert np.linalg.norm(a) > 0 and np.linalg.norm(b - 2.0 * a) <= 1e-14 * np.linalg.norm(a)
        assert gt.flops > 0 and gt.shuffle_tasks > 0
    finally:
        gt.close()


ONEDOT = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "guess1dot_*.npz")))


def make_onedot(rec, root, device):
    """One-dot record -> b2d_guess_desc: mode 1 (dot on the system side; "oldright" is the reference's newenvstateinfo) or mode 2."""
    p = "gw%d." % root
    transpose = int(rec["gw.nroots"][1]) != 0
    names = {"left": "left", "right": "right", "oldleft": "oldleft", "oldcol": "oldcol"}
    if transpose:
        names.update({"sys": "sys", "dot": "dot", "oldright": "newenv"})
    tabs = {k: {key[len(p + v + "."):]: rec[key] for key in rec if key.startswith(p + v + ".")} for k, v in names.items()}
    return hotpath.GuessTransform(rec[p + "dq"][:3], tabs, rec[p + "old.allowed"], rec[p + "lrot.shape"][:, 1], rec[p + "rrot.shape"][:, 1],
                                  device=device, mode=1 if transpose else 2)


@pytest.mark.parametrize("path", ONEDOT, ids=[os.path.basename(f)[:-4] for f in ONEDOT])
def test_planned_onedot_transform_reproduces_the_reference_trial_vector(path):
    """One-dot branch (GuessWave::onedot_transform_wavefunction, guess_wavefunction.C:832-936), both dot positions."""
    rec = dict(np.load(path))
    for root in range(int(rec["gw.nroots"][0])):
        p = "gw%d." % root
        gt = make_onedot(rec, root, device=-1)
        try:
            assert gt.trial_size == rec[p + "trial"].size
            got = execute_plan(gt, rec[p + "old.data"], rec[p + "lrot.data"], rec[p + "rrot.data"])
            ref = rec[p + "trial"]
            err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
            assert err < 1e-13, (path, root, err)
            assert (gt.shuffle_tasks > 0) == (int(rec["gw.nroots"][1]) != 0)       # the rotate-only mode has no shuffle
        finally:
            gt.close()


TRANSPOSE = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "guessT_*.npz")))


def make_transpose(rec, root, device):
    p = "gw%d." % root
    tabs = {k: {key[len(p + k + "."):]: rec[key] for key in rec if key.startswith(p + k + ".")} for k in ("left", "right", "oldleft", "oldcol")}
    return hotpath.GuessTransform(rec[p + "dq"][:3], tabs, rec[p + "old.allowed"], None, None, device=device, mode=3)


@pytest.mark.parametrize("path", TRANSPOSE, ids=[os.path.basename(f)[:-4] for f in TRANSPOSE])
def test_planned_transpose_guess_reproduces_the_reference_trial_vector(path):
    """First block iteration of a sweep (GuessWave::transpose_previous_wavefunction, guess_wavefunction.C:55-84): a signed
    transposition, bit-exact."""
    rec = dict(np.load(path))
    assert TRANSPOSE
    for root in range(int(rec["gw.nroots"][0])):
        p = "gw%d." % root
        gt = make_transpose(rec, root, device=-1)
        try:
            got = execute_plan(gt, rec[p + "old.data"], np.zeros(0), np.zeros(0))
            assert np.array_equal(got, rec[p + "trial"])
            assert gt.flops == 0 and gt.shuffle_rounds == 1
        finally:
            gt.close()


TRANSPOSE1 = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "guessT1_*.npz")))


def make_onedot_transpose(rec, root, device):
    p = "gw%d." % root
    tabs = {k: {key[len(p + k + "."):]: rec[key] for key in rec if key.startswith(p + k + ".")} for k in ("left", "sys", "dot", "right", "oldleft", "oldcol")}
    return hotpath.GuessTransform(rec[p + "dq"][:3], tabs, rec[p + "old.allowed"], None, None, device=device, mode=4)


@pytest.mark.parametrize("path", TRANSPOSE1, ids=[os.path.basename(f)[:-4] for f in TRANSPOSE1])
def test_planned_onedot_transpose_guess_reproduces_the_reference_trial_vector(path):
    """First block iteration of a one-dot sweep (GuessWave::onedot_transpose_wavefunction, guess_wavefunction.C:140-198)."""
    rec = dict(np.load(path))
    for root in range(int(rec["gw.nroots"][0])):
        p = "gw%d." % root
        gt = make_onedot_transpose(rec, root, device=-1)
        try:
            got = execute_plan(gt, rec[p + "old.data"], np.zeros(0), np.zeros(0))
            ref = rec[p + "trial"]
            assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-13
            assert gt.flops == 0 and gt.shuffle_tasks > 0
        finally:
            gt.close()


def test_planner_rejects_inconsistent_descriptions():
    """Host-side validation (the reference aborts on such inputs; the C ABI returns an error with a message)."""
    rec = dict(np.load(FIXTURES[0]))
    p = "gw0."
    tabs = tables(rec, 0)
    lcols, rcols = rec[p + "lrot.shape"][:, 1].copy(), rec[p + "rrot.shape"][:, 1].copy()
    with pytest.raises(hotpath.B2DError, match="mode must be"):
        _plan_with_mode(7)
    bad = lcols.copy()
    k = int(np.flatnonzero(bad > 0)[0])
    bad[k] += 1                                          # a left rotation matrix with one column too many
    with pytest.raises(hotpath.B2DError, match="left rotation matrix does not match"):
        hotpath.GuessTransform(rec[p + "dq"][:3], tabs, rec[p + "old.allowed"], bad, rcols, device=-1)
    broken = dict(tabs)
    broken["oldright"] = dict(tabs["oldright"])
    broken["oldright"]["unc.dims"] = tabs["oldright"]["unc.dims"] + 1      # pieces no longer add up to the collected sectors
    with pytest.raises(hotpath.B2DError, match="do not add up"):
        hotpath.GuessTransform(rec[p + "dq"][:3], broken, rec[p + "old.allowed"], lcols, rcols, device=-1)


def _plan_with_mode(mode):
    lib = hotpath._lib.load()
    import ctypes as C
    ctx = C.c_void_p()
    assert lib.b2d_create(-1, C.byref(ctx)) == 0
    try:
        d = hotpath._lib.GuessDescC()
        d.mode = mode
        out = np.zeros(8)
        rc = lib.b2d_guess_plan(ctx, C.byref(d), out.ctypes.data_as(hotpath._lib.c_f64p), 8)
        if rc:
            raise hotpath.B2DError(lib.b2d_last_error(ctx).decode())
    finally:
        lib.b2d_destroy(ctx)
