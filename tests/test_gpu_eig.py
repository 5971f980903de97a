"""GPU: the per-sector eigen-decomposition of the reduced density matrix (diagonalise_dm, rotationmat.C:258-279 -> dsyev_,
MatrixBLAS.C:381-414) for sectors of every size, through the C ABI: sectors of up to 64 states in the single-CTA Jacobi kernel,
larger ones in the block Jacobi kernel (block_b200/csrc/eig_block_jacobi.cuh).  The matrices are density-matrix-like: positive
semi-definite with a spectrum that decays over 20 orders of magnitude, so that the 1e-14 clamp and the 1e-13 keep threshold of the
reference (rotationmat.C:161,274) cut through the near-null space.  Checked against LAPACK (numpy.linalg.eigh = dsyevd):
eigenvalues to 1e-13 absolute (the bar the golden records use), orthonormal eigenvectors, residual |rho v - lambda v|, identical kept
counts and discarded weight from the oracle's restatement of sort_weights / assign_matrix_by_dm."""
import numpy as np
import pytest

from block_b200 import hotpath
from oracle import dmrg_oracle as O

HAM, OVERLAP = 0, 13


def tiny_block(dims):
    """A left child with the given sector sizes (N = 0, 1, 2, ... all spin 0: no operator couples them) and a one-sector right child."""
    nq = len(dims)
    q = np.array([[2 * i, 0, 0] for i in range(nq)], np.int32)
    eye = np.eye(nq, dtype=np.uint8)

    def ops(n, dd):
        a = np.eye(n, dtype=np.uint8)
        data = np.concatenate([np.eye(int(d)).ravel() for d in dd])
        return [hotpath.OperatorSpec(HAM, (), 0, (0, 0, 0), False, a, 0.0 * data), hotpath.OperatorSpec(OVERLAP, (), 0, (0, 0, 0), False, a, data)]
    left = hotpath.BlockSpec(q, np.asarray(dims, np.int32), sites=(0,), loop=True, ops=ops(nq, dims))
    rq = np.array([[2 * i, 0, 0] for i in range(nq)], np.int32)[::-1].copy()
    right = hotpath.BlockSpec(rq, np.ones(nq, np.int32), sites=(1,), loop=False, ops=ops(nq, np.ones(nq, np.int32)))
    del eye
    return left, right, (2 * (nq - 1), 0, 0)


def density_like(d, rng, decades=20.0):
    """rho = Q diag(w) Q^T with w log-uniform over `decades` orders of magnitude below 1 (plus exact zeros), trace 1."""
    qmat, _ = np.linalg.qr(rng.standard_normal((d, d)))
    w = 10.0 ** (-decades * np.sort(rng.random(d)))
    w[d - d // 8:] = 0.0
    w /= w.sum()
    rho = (qmat * w) @ qmat.T
    return 0.5 * (rho + rho.T)


@pytest.mark.gpu
@pytest.mark.parametrize("dims", [(65, 96, 3, 64), (130, 257, 40), (700, 33, 1)], ids=["65-96", "130-257", "700"])
def test_block_jacobi_against_lapack(dims):
    rng = np.random.default_rng(sum(dims))
    left, right, dq = tiny_block(dims)
    sb = hotpath.SpinBlock(left, right, dq, norbs=2, device=0)
    try:
        rho = [density_like(d, rng) for d in dims]
        rho = [r / len(dims) for r in rho]
        sb.set_density(rho)
        l0 = sb.kernel_launches()
        evals = sb.diagonalise_dm()
        assert sb.kernel_launches() - l0 >= 3
        ref = [np.linalg.eigvalsh(r) for r in rho]
        for a, b in zip(evals, ref):
            b = np.where(b < 1e-14, 0.0, b)
            assert np.abs(a - b).max() < 1e-13, np.abs(a - b).max()
        keep = sum(dims)
        kept, err, rot = sb.select_states(keep)
        ref_ev, ref_vec = O.diagonalise_dm(rho)
        ref_kept, ref_err = O.select_states(ref_ev, keep)
        ref_rot = O.rotation_matrices(ref_vec, ref_kept)
        assert list(kept) == [r.shape[1] for r in ref_rot]          # the 1e-13 keep threshold selects the same number of states per sector
        assert abs(err - ref_err) < 1e-12
        for q, u in enumerate(rot):
            k = u.shape[1]
            if k == 0:
                continue
            assert np.abs(u.T @ u - np.eye(k)).max() < 1e-12          # orthonormal
            lam = np.einsum("ik,ij,jk->k", u, rho[q], u)
            res = np.abs(rho[q] @ u - u * lam).max()
            assert res < 1e-14, res                                    # eigenvectors of rho (absolute, |rho| <= 1)
            # the well-separated part of the retained subspace is the reference's: compare projectors on the states above 1e-9
            big = lam > 1e-9
            pr = ref_rot[q][:, : int(big.sum())]
            assert np.abs(u[:, big] @ u[:, big].T - pr @ pr.T).max() < 1e-6
    finally:
        sb.close()


@pytest.mark.gpu
def test_block_jacobi_is_deterministic():
    dims = (200, 90)
    rng = np.random.default_rng(5)
    left, right, dq = tiny_block(dims)
    rho = [density_like(d, rng) / 2 for d in dims]
    out = []
    for _ in range(2):
        sb = hotpath.SpinBlock(left, right, dq, norbs=2, device=0)
        try:
            sb.set_density(rho)
            ev = sb.diagonalise_dm()
            kept, err, rot = sb.select_states(sum(dims))
            out.append((np.concatenate(ev), np.concatenate([r.ravel() for r in rot])))
        finally:
            sb.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
