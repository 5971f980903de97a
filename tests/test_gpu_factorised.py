"""GPU: FACTORISED enlarged-block operators (option "factorised"; SURVEY.md 7 "hard parts", operatorfunctions.C:188-250 rowstride /
colstride structure) against the materialised form, through the C ABI, on a synthetic big block whose two children are products
(renormalised block) x (one-site dot) built on the device from THEIR children (b2d_build_enlarged_op with the library's own planner and
random integrals): the same grandchildren, the same products - once written out as dense enlarged-block operators by the scatter kernel,
once kept as lists of scaled sub-blocks of the renormalised operators.  sigma, diag(H), the density matrix with perturbative noise and
the rotated operators must agree to rounding (1e-13 relative: only the summation order of the pre-summed factor blocks differs), while
the factorised form holds a fraction of the memory and executes a fraction of the flops (the structural zeros of the Kronecker blocks
are never touched).  The real-reference parity of the factorised form is pinned by tests/test_gpu_opbuild.py (every operator against
the reference's Op::build) and tests/test_gpu_dropin.py (whole sweeps)."""
import numpy as np
import pytest

from block_b200 import synthetic as S


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="module")
def pair():
    case = S.make_product_case(norbs=14, nelec=14, M=260, seed=11)
    mat = S.make_big_block_from_products(case, device=0, factorised=False)
    fac = S.make_big_block_from_products(case, device=0, factorised=True)
    yield case, mat, fac
    mat.close(); fac.close()


@pytest.mark.gpu
def test_sigma_identical(pair):
    case, mat, fac = pair
    assert mat.size == fac.size and mat.size > 10000
    assert abs(mat.sigma_flops() - fac.sigma_flops()) <= 1e-9 * mat.sigma_flops()      # ALGORITHMIC flops: the reference's dgemm count
    sm, sf = mat.plan_stats(), fac.plan_stats()
    assert sf["flops_executed"] < 0.75 * sm["flops_executed"]                              # structural zeros skipped
    assert sf["arena_doubles"] < sm["arena_doubles"]                                       # nothing materialised but the pre-summed factor blocks
    rng = np.random.default_rng(3)
    for _ in range(2):
        x = rng.standard_normal(mat.size)
        a, b = mat.multiplyH(x), fac.multiplyH(x)
        assert rel(b, a) < 1e-13, rel(b, a)
    # linearity through the factorised path and bit-reproducibility
    x, y = rng.standard_normal(mat.size), rng.standard_normal(mat.size)
    assert rel(fac.multiplyH(2.0 * x - 3.0 * y), 2.0 * fac.multiplyH(x) - 3.0 * fac.multiplyH(y)) < 1e-12
    assert np.array_equal(fac.multiplyH(x), fac.multiplyH(x))
    print("W = %d, executed flops factorised / materialised = %.3f, operator memory %.3f" %
          (mat.size, sf["flops_executed"] / sm["flops_executed"], sf["arena_doubles"] / sm["arena_doubles"]))


@pytest.mark.gpu
def test_diagonal_identical(pair):
    case, mat, fac = pair
    a, b = mat.diagonalH(), fac.diagonalH()
    assert rel(b, a) < 1e-13, rel(b, a)


@pytest.mark.gpu
def test_density_with_noise_and_rotation_identical(pair):
    case, mat, fac = pair
    rng = np.random.default_rng(5)
    psi = rng.standard_normal(mat.size)
    psi /= np.linalg.norm(psi)
    ra = mat.make_density([psi], [1.0], noise=1e-4)
    rb = fac.make_density([psi], [1.0], noise=1e-4)
    for x, y in zip(rb, ra):
        assert np.abs(x - y).max() < 1e-14 + 1e-12 * np.abs(y).max()
    # the same rotation matrices on both: random orthonormal columns keeping ~ half of every sector
    rot = []
    for d in mat.left.dims:
        d = int(d)
        k = max(1, d // 2)
        q, _ = np.linalg.qr(rng.standard_normal((d, d)))
        rot.append(np.ascontiguousarray(q[:, :k]))
    mat.set_rotation_matrices(rot); fac.set_rotation_matrices(rot)
    oa, da, opsa = mat.transform_operators()
    ob, db, opsb = fac.transform_operators()
    assert list(oa) == list(ob) and list(da) == list(db) and len(opsa) == len(opsb)
    worst = 0.0
    for (ma, xa), (mb, xb) in zip(opsa, opsb):
        assert np.array_equal(ma, mb)
        if xa.size:
            worst = max(worst, float(np.abs(xa - xb).max() / max(1.0, np.abs(xa).max())))
    assert worst < 1e-13, worst
