"""CPU restatement of the reference's construction of ENLARGED-BLOCK operators - the scatter the
reference runs inside every multiplyH in direct mode and again inside transform_operators, and the largest part of a small
sweep's wall time once the hot path is on the GPU (profiles/README.md, `host_op_build_s`).

TEST INFRASTRUCTURE ONLY (like everything under oracle/): nothing under block_b200/ imports it.  It is the checker a device
implementation of N2 will be held to; parity PINNED against records of the real reference (tests/golden/opbuild_*.npz, made by
tests/golden/make_opbuild_golden.py from oracle/_ref/block_dump): tests/test_opbuild_oracle.py.

Restated (file:line under the reference root):
    operatorfunctions::TensorTrace / TensorTraceElement        operatorfunctions.C:19-117
    operatorfunctions::TensorProduct / TensorProductElement    operatorfunctions.C:146-254
    MatrixTensorProduct                                        MatrixBLAS.C:125-200 (general Kronecker form; with -DFAST_MTP the
                                                               reference assumes the right factor is 1 x 1, which it is for a one-site dot)
    SparseMatrix::allocate                                     BaseOperator.C:123-145
    Cre::build, CreDes::build, CreCre::build, Overlap::build   Operators.C:453-487, 624-691, 847-900, 2597-2625
The complementary operators (CreDesComp, DesDesComp, CreCreDesComp: Operators.C:1059-2320) and Ham::build (:2322-2395) are sums of
the same two primitives weighted by one- and two-electron integrals; they are the next part to restate.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import dmrg_oracle as O

HAM, CRE, CRE_CRE, DES_DESCOMP, CRE_DES, CRE_DESCOMP, CRE_CRE_DESCOMP, OVERLAP = 0, 1, 2, 3, 4, 5, 6, 13


@dataclass
class ProductInfo:
    """StateInfo of an enlarged block c = left (x) right after CollectQuanta (StateInfo.h:113-147)."""
    q: np.ndarray            # collected quanta (nq, 3)
    dims: np.ndarray
    left: O.Block
    right: O.Block
    lmap: np.ndarray         # leftUnMapQuanta[uncollected index]
    rmap: np.ndarray         # rightUnMapQuanta[uncollected index]
    unc_dims: np.ndarray     # unCollectedStateInfo->quantaStates
    old_to_new: list         # oldToNewState[collected index] -> uncollected indices

    @staticmethod
    def from_record(rec, prefix="L.", left="LL.", right="LR."):
        from . import dumpio
        q = rec[prefix + "q"].astype(np.int64).reshape(-1, 3)
        begin = rec[prefix + "si.old_to_new_begin"]
        flat = rec[prefix + "si.old_to_new"]
        o2n = [[int(x) for x in flat[int(begin[k]):int(begin[k + 1])]] for k in range(len(begin) - 1)]
        assert int(rec[prefix + "si.left_is_LL"][0]) == 1
        return ProductInfo(q=q, dims=rec[prefix + "dims"].astype(np.int64), left=dumpio.block_from(rec, left), right=dumpio.block_from(rec, right),
                           lmap=rec[prefix + "si.lmap"].astype(np.int64), rmap=rec[prefix + "si.rmap"].astype(np.int64),
                           unc_dims=rec[prefix + "si.uncollected_dims"].astype(np.int64), old_to_new=o2n)


def allocate(pi: ProductInfo, optype, orbs, comp, dq, fermion) -> O.Op:
    """SparseMatrix::allocate(sr, sc) BaseOperator.C:123-145: block (i, j) exists iff q_i is in dq (+) q_j; zero-filled."""
    nq = len(pi.dims)
    allowed = np.zeros((nq, nq), bool)
    blocks = {}
    for i in range(nq):
        for j in range(nq):
            if O.qn_allow(tuple(pi.q[i]), tuple(dq), tuple(pi.q[j])):
                allowed[i, j] = True
                blocks[(i, j)] = np.zeros((int(pi.dims[i]), int(pi.dims[j])))
    return O.Op(optype=optype, orbs=tuple(orbs), comp=comp, dq=tuple(dq), fermion=fermion, allowed=allowed, blocks=blocks)


def _is_fermion(q) -> bool:
    return int(q[0]) % 2 != 0


def tensor_product(pi: ProductInfo, a: O.View, b: O.View, a_on_left: bool, c: O.Op, scale: float = 1.0):
    """operatorfunctions::TensorProduct(ablock, a, b, cblock, cstateinfo, c, scale) operatorfunctions.C:146-254: c += scale (a x b),
    `a` acting on the left child (a_on_left) or on the right child of the enlarged block."""
    if abs(scale) < 1e-20:
        return
    L, R = pi.left, pi.right
    cs = c.dq[1]
    for (cq, cqp), cel in c.blocks.items():
        row = 0
        for oi in pi.old_to_new[cq]:
            col = 0
            for oj in pi.old_to_new[cqp]:
                if a_on_left:
                    aq, aqp, bq, bqp = int(pi.lmap[oi]), int(pi.lmap[oj]), int(pi.rmap[oi]), int(pi.rmap[oj])
                else:
                    aq, aqp, bq, bqp = int(pi.rmap[oi]), int(pi.rmap[oj]), int(pi.lmap[oi]), int(pi.lmap[oj])
                if a.allowed(aq, aqp) and b.allowed(bq, bqp):
                    if a_on_left:                                                                   # :205-218
                        sb = O.ninej(L.q[aqp][1], R.q[bqp][1], pi.q[cqp][1], a.spin, b.spin, cs, L.q[aq][1], R.q[bq][1], pi.q[cq][1])
                        sb *= b.scaling(R.q[bq], R.q[bqp])
                        sa = scale * a.scaling(L.q[aq], L.q[aqp])
                        if b.fermion and _is_fermion(L.q[aqp]):
                            sb = -sb
                        blk = np.kron(sa * a.mat(aq, aqp), sb * b.mat(bq, bqp))
                    else:                                                                           # :219-236
                        sb = O.ninej(L.q[bqp][1], R.q[aqp][1], pi.q[cqp][1], b.spin, a.spin, cs, L.q[bq][1], R.q[aq][1], pi.q[cq][1])
                        sb *= b.scaling(L.q[bq], L.q[bqp])
                        sa = scale * a.scaling(R.q[aq], R.q[aqp])
                        if a.fermion and _is_fermion(L.q[bqp]):
                            sb = -sb
                        blk = np.kron(sb * b.mat(bq, bqp), sa * a.mat(aq, aqp))
                    cel[row:row + blk.shape[0], col:col + blk.shape[1]] += blk
                col += int(pi.unc_dims[oj])
            row += int(pi.unc_dims[oi])


def tensor_trace(pi: ProductInfo, a: O.View, a_on_left: bool, c: O.Op, scale: float = 1.0):
    """operatorfunctions::TensorTrace(ablock, a, cblock, cstateinfo, c, scale) operatorfunctions.C:19-117: c += scale (a x 1)."""
    if abs(scale) < 1e-20:
        return
    L, R = pi.left, pi.right
    cs = c.dq[1]
    for (cq, cqp), cel in c.blocks.items():
        row = 0
        for oi in pi.old_to_new[cq]:
            col = 0
            for oj in pi.old_to_new[cqp]:
                if a_on_left:
                    aq, aqp, bq, bqp = int(pi.lmap[oi]), int(pi.lmap[oj]), int(pi.rmap[oi]), int(pi.rmap[oj])
                    nb = int(R.dims[bq])
                else:
                    aq, aqp, bq, bqp = int(pi.rmap[oi]), int(pi.rmap[oj]), int(pi.lmap[oi]), int(pi.lmap[oj])
                    nb = int(L.dims[bq])
                if a.allowed(aq, aqp) and bq == bqp:
                    if a_on_left:                                                                   # :83-95
                        sb = O.ninej(L.q[aqp][1], R.q[bqp][1], pi.q[cqp][1], a.spin, 0, cs, L.q[aq][1], R.q[bq][1], pi.q[cq][1])
                        blk = np.kron(scale * a.mat(aq, aqp), sb * np.eye(nb))
                    else:                                                                           # :96-107
                        sb = O.ninej(L.q[bqp][1], R.q[aqp][1], pi.q[cqp][1], 0, a.spin, cs, L.q[bq][1], R.q[aq][1], pi.q[cq][1])
                        if a.fermion and _is_fermion(L.q[bqp]):
                            sb = -sb
                        blk = np.kron(sb * np.eye(nb), scale * a.mat(aq, aqp))
                    cel[row:row + blk.shape[0], col:col + blk.shape[1]] += blk
                col += int(pi.unc_dims[oj])
            row += int(pi.unc_dims[oi])


# ----------------------------------------------------------------------------------------------------------------------
# Op::build for the normal operators of an enlarged block
# ----------------------------------------------------------------------------------------------------------------------
def _find(block: O.Block, optype, orbs, dq=None):
    """get_op_rep(optype, deltaQuantum, i[, j]): the component whose deltaQuantum matches (op_components.h get_op_rep)."""
    for op in block.ops:
        if op.optype == optype and op.orbs == tuple(orbs) and (dq is None or tuple(op.dq) == tuple(dq)):
            return op
    return None


def _has(block: O.Block, optype, orbs) -> bool:
    return _find(block, optype, orbs) is not None


def _overlap(block: O.Block) -> O.View:
    return O.View(_find(block, OVERLAP, ()))


def _product_or_trace(pi, op, on_left, c):
    """`TensorTrace` if the other child has no sites, else `TensorProduct` with its OVERLAP (Operators.C:466-472, 640-648, 862-870)."""
    other = pi.right if on_left else pi.left
    if len(other.sites) == 0:
        tensor_trace(pi, O.View(op), on_left, c)
    elif on_left:
        tensor_product(pi, O.View(op), _overlap(other), True, c)       # TensorProduct(leftBlock, *op, *Overlap, ...)
    else:
        return False
    return True


def build_overlap(pi: ProductInfo) -> O.Op:
    """Overlap::build Operators.C:2597-2625."""
    c = allocate(pi, OVERLAP, (), 0, (0, 0, 0), False)
    if len(pi.right.sites) == 0:
        tensor_trace(pi, _overlap(pi.left), True, c)
    else:
        tensor_product(pi, _overlap(pi.right), _overlap(pi.left), False, c)    # TensorProduct(rightBlock, *op2, *op, ...)
    return c


def build_cre(pi: ProductInfo, i: int, dq) -> O.Op:
    """Cre::build Operators.C:453-487."""
    c = allocate(pi, CRE, (i,), 0, dq, True)
    L, R = pi.left, pi.right
    if _has(L, CRE, (i,)):
        _product_or_trace(pi, _find(L, CRE, (i,), dq), True, c)
    elif _has(R, CRE, (i,)):
        tensor_product(pi, _overlap(L), O.View(_find(R, CRE, (i,), dq)), True, c)   # TensorProduct(leftBlock, *Overlap, *op, ...)
    else:
        raise ValueError("Cre::build: orbital %d on neither child" % i)
    return c


def build_credes(pi: ProductInfo, i: int, j: int, comp: int, dq) -> O.Op:
    """CreDes::build Operators.C:624-691 (blocks without explicit DES operators: the Transposeview branches)."""
    c = allocate(pi, CRE_DES, (i, j), comp, dq, False)
    L, R = pi.left, pi.right
    if _has(L, CRE_DES, (i, j)):
        _product_or_trace(pi, _find(L, CRE_DES, (i, j), dq), True, c)
    elif _has(R, CRE_DES, (i, j)):
        tensor_product(pi, O.View(_find(R, CRE_DES, (i, j), dq)), _overlap(L), False, c)   # TensorProduct(rightBlock, *op, *Overlap, ...)
    elif _has(L, CRE, (i,)):
        op1, op2 = _find(L, CRE, (i,)), _find(R, CRE, (j,))
        tensor_product(pi, O.View(op1), O.View(op2, True), True, c, 1.0)
    elif _has(R, CRE, (i,)):
        op1, op2 = _find(R, CRE, (i,)), _find(L, CRE, (j,))
        parity = O.commute_parity(tuple(op1.dq), O.neg(tuple(op2.dq)), tuple(dq))
        tensor_product(pi, O.View(op1), O.View(op2, True), False, c, parity)
    else:
        raise ValueError("CreDes::build: orbitals %d %d not available" % (i, j))
    return c


def build_crecre(pi: ProductInfo, i: int, j: int, comp: int, dq) -> O.Op:
    """CreCre::build Operators.C:847-900."""
    c = allocate(pi, CRE_CRE, (i, j), comp, dq, False)
    L, R = pi.left, pi.right
    if _has(L, CRE_CRE, (i, j)):
        _product_or_trace(pi, _find(L, CRE_CRE, (i, j), dq), True, c)
    elif _has(R, CRE_CRE, (i, j)):
        tensor_product(pi, _overlap(L), O.View(_find(R, CRE_CRE, (i, j), dq)), True, c)    # TensorProduct(leftBlock, *Overlap, *op, ...)
    elif _has(L, CRE, (i,)):
        tensor_product(pi, O.View(_find(L, CRE, (i,))), O.View(_find(R, CRE, (j,))), True, c, 1.0)
    elif _has(R, CRE, (i,)):
        op1, op2 = _find(R, CRE, (i,)), _find(L, CRE, (j,))
        parity = O.commute_parity(tuple(op1.dq), tuple(op2.dq), tuple(dq))
        tensor_product(pi, O.View(op1), O.View(op2), False, c, parity)
    else:
        raise ValueError("CreCre::build: orbitals %d %d not available" % (i, j))
    return c


def build_normal_operator(pi: ProductInfo, ref_op: O.Op) -> O.Op:
    """Build the enlarged-block operator that corresponds to `ref_op` (type, orbitals, component, deltaQuantum)."""
    if ref_op.optype == CRE:
        return build_cre(pi, ref_op.orbs[0], ref_op.dq)
    if ref_op.optype == CRE_DES:
        return build_credes(pi, ref_op.orbs[0], ref_op.orbs[1], ref_op.comp, ref_op.dq)
    if ref_op.optype == CRE_CRE:
        return build_crecre(pi, ref_op.orbs[0], ref_op.orbs[1], ref_op.comp, ref_op.dq)
    if ref_op.optype == OVERLAP:
        return build_overlap(pi)
    raise ValueError("operator type %d is not restated yet" % ref_op.optype)
