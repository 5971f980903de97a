"""Parity of the CUDA hot path (through the C ABI) with the oracle and with records of the real reference.
Tolerances are north_star's: sigma <= 1e-10 relative, energies <= 1e-8 Eh, identical retained sectors / counts."""
import numpy as np
import pytest

from block_b200 import hotpath
from oracle import dmrg_oracle as O
from oracle import dumpio

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.fixture(scope="module")
def gpu_block(golden):
    rec, big = golden
    sb = hotpath.spinblock_from_record(rec, device=0)
    yield rec, big, sb
    sb.close()


def test_operator_roundtrip(gpu_block):
    rec, big, sb = gpu_block
    for side, blk in enumerate((sb.left, sb.right)):
        for k in (0, len(blk.ops) // 2, len(blk.ops) - 1):
            assert np.array_equal(sb.download_op(side, sb.op_ids[side][k]), blk.ops[k].data)    # bit-exact


def test_wavefunction_roundtrip(gpu_block):
    rec, big, sb = gpu_block
    sb.upload(0, rec["rpsi"])
    assert np.array_equal(sb.download(0), rec["rpsi"])


def test_sigma_matches_reference(gpu_block):
    rec, big, sb = gpu_block
    v = sb.multiplyH(rec["rpsi"])
    assert rel(v, rec["rsigma"]) < 1e-10
    assert rel(v, rec["rsigma"]) < 1e-13          # what FP64 DMMA actually delivers
    for i in range(int(rec["meta"][4])):
        assert rel(sb.multiplyH(rec["psi%d" % i]), rec["sigma%d" % i]) < 1e-10


def test_context_reuse_across_block_iterations(golden):
    """b2d_reset: one context serves a sequence of different big blocks (what the drop-in sweep does); results must be those of a
    fresh context, bit for bit."""
    import glob
    import os
    rec, big = golden
    others = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*_site*.npz")))
    other = dumpio.read_records(others[0])
    sb = hotpath.spinblock_from_record(other, device=0)       # start on ANOTHER block, run something, then reset onto `rec`
    try:
        sb.multiplyH(other["rpsi"])
        nro = int(other["meta"][4])
        sb.RenormaliseFrom([other["guess%d" % i] for i in range(nro)], other["weights"], float(other["dav_tol"][0]), int(other["meta"][5]))
        sb.transform_operators()
        L, R = hotpath.block_spec_from_record(rec, "L."), hotpath.block_spec_from_record(rec, "R.")
        norbs = len(rec["spin_orbs_symmetry"]) // 2
        sb.reset(L, R, tuple(int(x) for x in rec["psi_dq"]), core_energy=float(rec["meta_f"][3]), hubbard=int(rec["meta"][7]) == hotpath.HUBBARD, norbs=norbs)
        v = sb.multiplyH(rec["rpsi"])
        assert rel(v, rec["rsigma"]) < 1e-13
        fresh = hotpath.spinblock_from_record(rec, device=0)
        try:
            assert np.array_equal(v, fresh.multiplyH(rec["rpsi"]))
            nroots = int(rec["meta"][4])
            args = ([rec["guess%d" % i] for i in range(nroots)], rec["weights"], float(rec["dav_tol"][0]), int(rec["meta"][5]))
            a, b = sb.RenormaliseFrom(*args, noise=float(rec["rdm.args"][0])), fresh.RenormaliseFrom(*args, noise=float(rec["rdm.args"][0]))
            assert np.array_equal(a["energies"], b["energies"]) and list(a["kept"]) == list(b["kept"]) and a["error"] == b["error"]
        finally:
            fresh.close()
    finally:
        sb.close()


def test_sigma_accumulates_like_reference(gpu_block):
    rec, big, sb = gpu_block
    v0 = np.cos(np.arange(sb.size))
    v = sb.multiplyH(rec["rpsi"], v0.copy())
    assert rel(v - v0, rec["rsigma"]) < 1e-10


@pytest.mark.parametrize("cls", [0, 1, 2, 3])
def test_sigma_every_tile_class(golden, cls):
    rec, big = golden
    sb = hotpath.spinblock_from_record(rec, device=0, options={"tile_class": cls})
    try:
        assert rel(sb.multiplyH(rec["rpsi"]), rec["rsigma"]) < 1e-12
    finally:
        sb.close()


def test_sigma_chunked_workspace(golden):
    rec, big = golden
    sb = hotpath.spinblock_from_record(rec, device=0, options={"workspace_mb": 0.01})
    try:
        assert sb.plan_stats()["chunks"] > 1
        assert rel(sb.multiplyH(rec["rpsi"]), rec["rsigma"]) < 1e-12
    finally:
        sb.close()


@pytest.mark.parametrize("iters", [0, 1, 3])
def test_sigma_split_k_slices(golden, iters):
    """step-2 split-K (partial copies summed in a fixed order) must not change sigma; 0 disables the split."""
    rec, big = golden
    sb = hotpath.spinblock_from_record(rec, device=0, options={"slice_iters": iters})
    try:
        v1 = sb.multiplyH(rec["rpsi"])
        assert rel(v1, rec["rsigma"]) < 1e-12
        assert np.array_equal(v1, sb.multiplyH(rec["rpsi"]))          # deterministic run to run
        v0 = np.sin(np.arange(sb.size))
        assert rel(sb.multiplyH(rec["rpsi"], v0.copy()) - v0, rec["rsigma"]) < 1e-10
    finally:
        sb.close()


def test_tensor_multiply_single_terms(gpu_block):
    rec, big, sb = gpu_block
    terms = O.h_terms(big)
    lo, ro, fl, sc, ow = sb.terms(all_ranks=True)
    c = big.unflatten(rec["rpsi"])
    for k in sorted(set([0, 1, 2, len(terms) // 2, len(terms) - 1])):
        v = big.zeros()
        O.tensor_multiply(big, terms[k][0], terms[k][1], c, v, 0, terms[k][2])
        ref = big.flatten(v)
        got = sb.TensorMultiply(int(lo[k]), int(ro[k]), rec["rpsi"], np.zeros(sb.size), bool(fl[k] & 1), bool(fl[k] & 2), 0, float(sc[k]))
        assert np.linalg.norm(got - ref) <= 1e-12 * max(np.linalg.norm(ref), 1e-30)


def test_tensor_multiply_one_operator_form(gpu_block):
    """operatorfunctions.C:331-404 on both children, plain and Transposeview, into the dQ-shifted target sectors."""
    rec, big, sb = gpu_block
    wl = O.WaveLayout(big, big.psi_dq)
    c = wl.unflatten(rec["rpsi"])
    checked = 0
    for side, blk in enumerate((big.left, big.right)):
        picks = {}
        for k, op in enumerate(blk.ops):
            picks.setdefault(op.optype, k)                      # first operator of every type
        for k in picks.values():
            op = blk.ops[k]
            for sign, t in ((+1, False), (-1, True)):
                for spin in range(abs(big.psi_dq[1] - op.dq[1]), big.psi_dq[1] + op.dq[1] + 1, 2):
                    dq = (big.psi_dq[0] + sign * op.dq[0], spin, O.irrep_mul(big.psi_dq[2], op.dq[2]))
                    vl = O.WaveLayout(big, dq)
                    assert sb.wavefunction_size(dq) == vl.size
                    if vl.size == 0:
                        continue
                    v = vl.zeros()
                    O.tensor_multiply_one(big, O.View(op, t), side == 0, c, wl, v, vl, 0.75)
                    ref = vl.flatten(v)
                    v0 = np.cos(np.arange(vl.size))
                    got = sb.TensorMultiplyOne(side, sb.op_ids[side][k], rec["rpsi"], v0.copy(), dq, transposed=t, scale=0.75)
                    assert np.linalg.norm(got - v0 - ref) <= 1e-12 * max(np.linalg.norm(ref), 1e-30)
                    checked += 1
    assert checked >= 6


def test_diagonal_matches_reference(gpu_block):
    rec, big, sb = gpu_block
    assert rel(sb.diagonalH(), rec["diag"]) < 1e-13


def test_davidson_matches_reference(gpu_block):
    rec, big, sb = gpu_block
    nroots = int(rec["meta"][4])
    ev, vecs, nmult = sb.block_davidson([rec["guess%d" % i] for i in range(nroots)], rec["diag"], float(rec["dav_tol"][0]),
                                        int(rec["dav_in"][4]), int(rec["dav_in"][5]))
    assert np.abs(ev - rec["dav_evals"][:nroots]).max() < 1e-8          # north_star: energies within 1e-8 Eh
    assert np.abs(ev - rec["dav_evals"][:nroots]).max() < 1e-10
    assert nmult == int(rec["dav_out"][0])                              # same number of H applications
    for i in range(nroots):
        assert abs(abs(np.dot(vecs[i], rec["psi%d" % i])) - 1.0) < 1e-8


def test_davidson_with_lower_states_matches_reference(gpu_block):
    """b2d_davidson_lower against the reference's own state-specific solve recorded in the fixture (one lower state)."""
    rec, big, sb = gpu_block
    ev, vecs, nmult = sb.block_davidson([rec["ss_guess"]], rec["diag"], float(rec["dav_tol"][0]), int(rec["dav_in"][4]), int(rec["dav_in"][5]),
                                        lowerStates=[rec["ss_lower"]])
    assert abs(ev[0] - rec["ss_eval"][0]) < 1e-8            # north_star: energies within 1e-8 Eh
    assert abs(ev[0] - rec["ss_eval"][0]) < 1e-10
    assert nmult == int(rec["ss_nmult"][0])                 # same number of H applications
    assert abs(abs(np.dot(vecs[0], rec["ss_psi"])) - 1.0) < 1e-8
    assert abs(np.dot(vecs[0], rec["ss_lower"])) / np.linalg.norm(rec["ss_lower"]) < 1e-9


def test_density_matches_reference(gpu_block):
    rec, big, sb = gpu_block
    nroots = int(rec["meta"][4])
    rho = sb.make_density([rec["psi%d" % i] for i in range(nroots)], rec["weights"], noise=float(rec["rdm.args"][0]))
    assert rel(np.concatenate([r.ravel() for r in rho]), rec["rdm.data"]) < 1e-13      # incl. add_onedot_noise in the *_noise fixtures


def test_truncation_identical_sectors_and_counts(gpu_block):
    rec, big, sb = gpu_block
    nroots = int(rec["meta"][4])
    sb.make_density([rec["psi%d" % i] for i in range(nroots)], rec["weights"], noise=float(rec["rdm.args"][0]))
    evals = sb.diagonalise_dm()
    rho = O.make_density_with_noise(big, [big.unflatten(rec["psi%d" % i]) for i in range(nroots)], rec["weights"], float(rec["rdm.args"][0]))
    ref_evals, _ = O.diagonalise_dm(rho)
    for a, b in zip(evals, ref_evals):
        assert np.abs(a - b).max() < 1e-13
    kept, err, rot = sb.select_states(int(rec["meta"][5]))
    ref_rot = dumpio.rotation_from(rec)
    assert list(kept) == [r.shape[1] for r in ref_rot]                  # bit-exact integer contract
    assert abs(err - rec["error"][0]) < 1e-12
    for q in range(len(rot)):
        if rot[q].shape[1]:
            assert np.abs(rot[q].T @ rot[q] - np.eye(rot[q].shape[1])).max() < 1e-12
            assert np.abs(rot[q] @ rot[q].T - ref_rot[q] @ ref_rot[q].T).max() < 1e-7


def test_truncation_with_cusolver_sectors(golden):
    """large sectors go through cusolverDnDsyevd instead of the Jacobi kernel: force it for every sector with > 2 states."""
    rec, big = golden
    sb = hotpath.spinblock_from_record(rec, device=0, options={"eig_jacobi_max": 2})
    try:
        nroots = int(rec["meta"][4])
        noise = float(rec["rdm.args"][0])
        sb.make_density([rec["psi%d" % i] for i in range(nroots)], rec["weights"], noise=noise)
        evals = sb.diagonalise_dm()
        rho = O.make_density_with_noise(big, [big.unflatten(rec["psi%d" % i]) for i in range(nroots)], rec["weights"], noise)
        ref_evals, _ = O.diagonalise_dm(rho)
        for a, b in zip(evals, ref_evals):
            assert np.abs(a - b).max() < 1e-13
        kept, err, rot = sb.select_states(int(rec["meta"][5]))
        ref_rot = dumpio.rotation_from(rec)
        assert list(kept) == [r.shape[1] for r in ref_rot]
        assert abs(err - rec["error"][0]) < 1e-12
        for q in range(len(rot)):
            if rot[q].shape[1]:
                assert np.abs(rot[q].T @ rot[q] - np.eye(rot[q].shape[1])).max() < 1e-12
                assert np.abs(rot[q] @ rot[q].T - ref_rot[q] @ ref_rot[q].T).max() < 1e-7
    finally:
        sb.close()


def test_transform_operators_matches_reference(gpu_block):
    rec, big, sb = gpu_block
    ref_rot = dumpio.rotation_from(rec)
    sb.set_rotation_matrices(ref_rot)
    old, dims, ops = sb.transform_operators()
    N = dumpio.block_from(rec, "N.")
    keepq = [q for q in range(len(ref_rot)) if ref_rot[q].shape[1] > 0]
    assert list(old) == keepq and list(dims) == list(N.dims)
    checked = 0
    for nop in N.ops:
        k = next((i for i, o in enumerate(big.left.ops) if o.optype == nop.optype and o.orbs == nop.orbs and o.comp == nop.comp), None)
        if k is None:
            continue
        allowed, data = ops[k]
        assert (allowed == nop.allowed).all()
        ref = np.concatenate([nop.blocks[(a, b)].ravel() for a in range(len(dims)) for b in range(len(dims)) if nop.allowed[a, b]] + [np.zeros(0)])
        assert np.abs(data - ref).max() < 1e-12
        checked += 1
    assert checked > 0


def test_renormalise_from_end_to_end(gpu_block):
    rec, big, sb = gpu_block
    nroots = int(rec["meta"][4])
    out = sb.RenormaliseFrom([rec["guess%d" % i] for i in range(nroots)], rec["weights"], float(rec["dav_tol"][0]), int(rec["meta"][5]),
                             int(rec["dav_in"][4]), int(rec["dav_in"][5]), noise=float(rec["rdm.args"][0]))
    assert np.abs(out["energies"] - rec["energies"][:nroots]).max() < 1e-8
    ref_rot = dumpio.rotation_from(rec)
    assert list(out["kept"]) == [r.shape[1] for r in ref_rot]
    assert abs(out["error"] - rec["error"][0]) < 1e-9
    assert out["n_multiply"] == int(rec["dav_out"][0])


def test_sigma_properties_linearity_and_symmetry(gpu_block):
    """size-independent properties used again at benchmark scale: H(ax+by) = aHx + bHy and <x|Hy> = <y|Hx>."""
    rec, big, sb = gpu_block
    rng = np.random.default_rng(7)
    x, y = rng.standard_normal(sb.size), rng.standard_normal(sb.size)
    hx, hy = sb.multiplyH(x), sb.multiplyH(y)
    assert rel(sb.multiplyH(2.0 * x - 3.0 * y), 2.0 * hx - 3.0 * hy) < 1e-12
    assert abs(np.dot(x, hy) - np.dot(y, hx)) < 1e-10 * np.linalg.norm(hx) * np.linalg.norm(y)
