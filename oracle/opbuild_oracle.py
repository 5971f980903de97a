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
    TensorOp (spin-coupled operator strings)                   tensor_operator.h:27-289
    SparseMatrix::calcCompfactor                               Operators.C:87-137
    CreDesComp::build, DesDesComp::build                       Operators.C:1059-1194, 1472-1616 (integral-weighted child products)
    CreCreDesComp::build -> opxop::cxcdcomp / dxcccomp         Operators.C:1905-1966, opxop.C:376-532
    Ham::build -> opxop::cxcddcomp / cdxcdcomp / ddxcccomp     Operators.C:2322-2395, opxop.C:22-143
i.e. EVERY operator type an energy sweep carries on an enlarged block.
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


# ----------------------------------------------------------------------------------------------------------------------
# complementary operators: spin-coupled operator strings (TensorOp, tensor_operator.h) contracted with the integrals
# (SparseMatrix::calcCompfactor, Operators.C:87-137), then the same TensorProduct as above with that scalar.
# Spin-adapted, abelian point group (irreps are bit patterns, product = XOR, one row per irrep).
# ----------------------------------------------------------------------------------------------------------------------
class Integrals:
    """v_1 / v_2 as the reference's accessors return them for SPIN-orbital indices (IntegralMatrix.C:32-46, 312-333, rhf): zero unless
    the spins of (i,k) and of (j,l) match, else the spatial value.  Spatial arrays come from the fixture (reordered orbitals)."""

    def __init__(self, v1, v2, orbital_irreps, screen_tol=(1e-20, 1e-20)):
        self.h1, self.h2 = np.asarray(v1), np.asarray(v2)
        self.irreps = [int(x) for x in orbital_irreps]       # per spatial orbital
        self.one_tol, self.two_tol = float(screen_tol[0]), float(screen_tol[1])

    def v1(self, i, j):
        return 0.0 if (i & 1) != (j & 1) else float(self.h1[i // 2, j // 2])

    def v2(self, i, j, k, l):
        if (i & 1) != (k & 1) or (j & 1) != (l & 1):
            return 0.0
        return float(self.h2[i // 2, j // 2, k // 2, l // 2])

    @staticmethod
    def from_record(rec):
        sos = rec["spin_orbs_symmetry"]
        return Integrals(rec["v1"], rec["v2"], [int(sos[2 * i]) for i in range(len(sos) // 2)], rec["screen_tol"])


class TensorOp:
    """tensor_operator.h:27-354: a spin tensor operator as coefficient vectors (one per Sz component, index (Spin - sz)/2) over
    products of spin-orbital creation (+1) / destruction (-1) operators."""

    def __init__(self, k=None, sign=None, ints: Integrals = None):
        self.empty = True
        self.spin, self.irrep, self.szops, self.opindices, self.optypes = 0, 0, [], [], []
        if k is None:
            return
        K = 2 * k                                                       # spatial_to_spin
        ind = [K + 1, K + 0] if sign < 0 else [K + 0, K + 1]            # :58-66
        self.empty = False
        self.spin = 1
        self.irrep = ints.irreps[k]                                     # abelian: -irrep == irrep
        self.optypes = [sign]
        self.szops = [[sign * 1.0, 0.0], [0.0, 1.0]]                    # :95-99
        self.opindices = [[ind[0]], [ind[1]]]

    def dn(self):
        return sum(self.optypes)

    def product(self, other: "TensorOp", pspin: int, pirrep: int) -> "TensorOp":
        """TensorOp::product tensor_operator.h:196-289 (`identical` is forced to false there)."""
        out = TensorOp()
        if pspin < abs(self.spin - other.spin) or pspin > self.spin + other.spin:
            raise ValueError("cannot combine spins %d and %d to %d" % (self.spin, other.spin, pspin))
        if (self.irrep ^ other.irrep) != pirrep:
            return out                                                  # empty
        out.empty = False
        out.optypes = self.optypes + other.optypes
        out.opindices = [a + b for a in self.opindices for b in other.opindices]
        n2 = len(other.opindices)
        out.szops = [[0.0] * len(out.opindices) for _ in range(pspin + 1)]
        for sz in range(pspin, -pspin - 1, -2):
            dst = out.szops[(pspin - sz) // 2]
            for sz1 in range(self.spin, -self.spin - 1, -2):
                for sz2 in range(other.spin, -other.spin - 1, -2):
                    cleb = O.clebsch(self.spin, sz1, other.spin, sz2, pspin, sz)
                    if abs(cleb) <= 1e-14:
                        continue
                    v1, v2 = self.szops[(self.spin - sz1) // 2], other.szops[(other.spin - sz2) // 2]
                    for i, a in enumerate(v1):
                        for j, b in enumerate(v2):
                            dst[i * n2 + j] += cleb * a * b
        out.spin, out.irrep = pspin, pirrep
        return out


def calc_compfactor(op1: TensorOp, op2: TensorOp, comp: str, ints: Integrals) -> float:
    """SparseMatrix::calcCompfactor(op1, op2, comp, v_2, integralIndex) Operators.C:87-137 (comp in CD, DD, CCD, CDD, C)."""
    factor = 0.0
    c1 = op1.szops[0]
    for sz2 in range(-op2.spin, op2.spin + 1, 2):
        c2 = op2.szops[(op2.spin - sz2) // 2]
        cleb = O.clebsch(op1.spin, op1.spin, op2.spin, sz2, 0, 0)
        if (op1.irrep ^ op2.irrep) != 0:
            cleb = 0.0
        if abs(cleb) <= 1e-14:
            continue
        for i1, a in enumerate(c1):
            for i2, b in enumerate(c2):
                if a == 0.0 or b == 0.0:
                    continue
                I1, I2 = op1.opindices[i1], op2.opindices[i2]
                if comp == "CD":
                    t = 0.5 * (-ints.v2(I1[0], I2[0], I2[1], I1[1]) - ints.v2(I2[0], I1[0], I1[1], I2[1])
                               + ints.v2(I2[0], I1[0], I2[1], I1[1]) + ints.v2(I1[0], I2[0], I1[1], I2[1]))
                elif comp == "DD":
                    t = 0.5 * ints.v2(I1[0], I1[1], I2[1], I2[0])
                elif comp == "CCD":
                    t = 0.5 * (ints.v2(I1[0], I1[1], I2[0], I1[2]) - ints.v2(I1[1], I1[0], I2[0], I1[2]))
                elif comp == "CDD":
                    t = 0.5 * (ints.v2(I2[0], I1[0], I1[2], I1[1]) - ints.v2(I1[0], I2[0], I1[2], I1[1]))
                elif comp == "C":
                    t = 0.5 * ints.v1(I1[0], I2[0])
                else:
                    raise ValueError(comp)
                factor += t * a * b / cleb
        break      # `found`: only the first Sz component with a non-vanishing coupling is used
    return factor


def _carry_over(pi, optype, i_j, dq, c):
    """The part of a complementary operator that already lives on one child: comp(L) x 1 and 1 x comp(R) (Operators.C:1076-1101)."""
    L, R = pi.left, pi.right
    if _has(L, optype, i_j):
        _product_or_trace(pi, _find(L, optype, i_j, dq), True, c)
    if len(R.sites) == 0:
        return False
    if _has(R, optype, i_j):
        tensor_product(pi, _overlap(L), O.View(_find(R, optype, i_j, dq)), True, c)
    return True


def build_credescomp(pi: ProductInfo, i: int, j: int, comp: int, dq, ints: Integrals) -> O.Op:
    """CreDesComp::build Operators.C:1059-1194 (non-BCS, blocks without explicit DES operators)."""
    c = allocate(pi, CRE_DESCOMP, (i, j), comp, dq, False)
    spin, sym = dq[1], dq[2]
    CD1 = TensorOp(i, 1, ints).product(TensorOp(j, -1, ints), spin, sym)          # the operator to be complemented
    if not _carry_over(pi, CRE_DESCOMP, (i, j), dq, c):
        return c
    L, R = pi.left, pi.right
    for k in L.sites:
        for l in R.sites:
            have = _has(L, CRE, (k,)) and _has(R, CRE, (l,))
            CD2 = TensorOp(k, 1, ints).product(TensorOp(l, -1, ints), spin, sym)
            if not CD2.empty:
                s = calc_compfactor(CD1, CD2, "CD", ints)
                if have and abs(s) > ints.two_tol:                                        # c+_k(L) x d_l(R)
                    tensor_product(pi, O.View(_find(L, CRE, (k,))), O.View(_find(R, CRE, (l,)), True), True, c, s)
            CD2 = TensorOp(l, 1, ints).product(TensorOp(k, -1, ints), spin, sym)
            if not CD2.empty:
                s = calc_compfactor(CD1, CD2, "CD", ints)
                if have and abs(s) > ints.two_tol:                                        # c+_l(R) x d_k(L)
                    op1, op2 = _find(R, CRE, (l,)), _find(L, CRE, (k,))
                    parity = O.commute_parity(tuple(op1.dq), O.neg(tuple(op2.dq)), tuple(dq))
                    tensor_product(pi, O.View(op1), O.View(op2, True), False, c, s * parity)
    return c


def build_desdescomp(pi: ProductInfo, i: int, j: int, comp: int, dq, ints: Integrals) -> O.Op:
    """DesDesComp::build Operators.C:1472-1616 (non-BCS, blocks without explicit DES operators)."""
    c = allocate(pi, DES_DESCOMP, (i, j), comp, dq, False)
    spin, sym = dq[1], dq[2]
    CC1 = TensorOp(i, 1, ints).product(TensorOp(j, 1, ints), spin, sym)
    if not _carry_over(pi, DES_DESCOMP, (i, j), dq, c):
        return c
    L, R = pi.left, pi.right
    for k in L.sites:
        for l in R.sites:
            DD2 = TensorOp(k, -1, ints).product(TensorOp(l, -1, ints), spin, sym)
            if DD2.empty:
                continue
            s = calc_compfactor(CC1, DD2, "DD", ints)
            s2 = calc_compfactor(CC1, TensorOp(l, -1, ints).product(TensorOp(k, -1, ints), spin, sym), "DD", ints)
            if _has(L, CRE, (k,)) and _has(R, CRE, (l,)) and abs(s) + abs(s2) > ints.two_tol:
                op1, op2 = _find(L, CRE, (k,)), _find(R, CRE, (l,))
                parity = O.commute_parity(O.neg(tuple(op1.dq)), O.neg(tuple(op2.dq)), tuple(dq))
                s += parity * s2
                if abs(s) > ints.two_tol:
                    tensor_product(pi, O.View(op1, True), O.View(op2, True), True, c, s)
    return c


def build_operator(pi: ProductInfo, ref_op: O.Op, ints: Integrals = None, hubbard=False) -> O.Op:
    """Any restated operator type of the enlarged block."""
    if ref_op.optype == CRE_CRE_DESCOMP:
        return build_crecredescomp(pi, ref_op.orbs[0], ref_op.dq, ints, hubbard)
    if ref_op.optype == HAM:
        return build_ham(pi, hubbard)
    if ref_op.optype == CRE_DESCOMP:
        return build_credescomp(pi, ref_op.orbs[0], ref_op.orbs[1], ref_op.comp, ref_op.dq, ints)
    if ref_op.optype == DES_DESCOMP:
        return build_desdescomp(pi, ref_op.orbs[0], ref_op.orbs[1], ref_op.comp, ref_op.dq, ints)
    return build_normal_operator(pi, ref_op)


# ----------------------------------------------------------------------------------------------------------------------
# three-index complementary operator and the Hamiltonian of the enlarged block: sums over products of one child's normal
# operators with the other child's two- / three-index complementary operators (no integrals appear explicitly here)
# ----------------------------------------------------------------------------------------------------------------------
def _recoupling(j2, j1, j21, phase_twice):
    """pow(-1, int(phase/2)) * sixj(j2, j1, j21, 1, 0, j2) * sqrt((j21+1)(j2+1))   (opxop.C:395, 421, 471, 504)."""
    return (-1.0) ** int(phase_twice / 2) * O.six_j(j2, j1, j21, 1, 0, j2) * np.sqrt((j21 + 1.0) * (j2 + 1.0))


def _comps(block: O.Block, optype, orbs):
    return [op for op in block.ops if op.optype == optype and op.orbs == tuple(orbs)]


def _cxcdcomp(pi, other_is_left, op1, I, c, scale):
    """opxop::cxcdcomp (operator form) opxop.C:376-445: c += c+_j (loop side) x CDcomp_{jI} or its transpose (other side)."""
    other = pi.left if other_is_left else pi.right
    j = op1.orbs[0]
    j1, j21 = op1.dq[1], c.dq[1]
    if j >= I:
        for op2 in _comps(other, CRE_DESCOMP, (j, I)):
            j2 = op2.dq[1]
            f = _recoupling(j2, j1, j21, 2 + j2)
            if not other_is_left:
                f *= O.commute_parity(tuple(op1.dq), tuple(op2.dq), tuple(c.dq))
            tensor_product(pi, O.View(op2), O.View(op1), other_is_left, c, f * scale)
    else:
        for op2 in _comps(other, CRE_DESCOMP, (I, j)):
            j2 = op2.dq[1]
            f = _recoupling(j2, j1, j21, 1 + 1 + 0 + j2)
            if not other_is_left:
                f *= O.commute_parity(tuple(op1.dq), O.neg(tuple(op2.dq)), tuple(c.dq))
            f *= -1.0 if j2 == 2 else 1.0                                  # TensorOp::getTransposeFactorCD, abelian
            tensor_product(pi, O.View(op2, True), O.View(op1), other_is_left, c, f * scale)


def _dxcccomp(pi, other_is_left, op1, K, c, scale, ints: Integrals):
    """opxop::dxcccomp (operator form, blocks without DES operators) opxop.C:447-532: c += d_j (loop side) x DDcomp_{kj}^T (other side)."""
    other = pi.left if other_is_left else pi.right
    k, i, transpose = K, op1.orbs[0], False
    if k < i:
        k, i, transpose = i, K, True
    iq, kq = (1, 1, ints.irreps[i]), (1, 1, ints.irreps[k])
    j1, j21 = op1.dq[1], c.dq[1]
    for op2 in _comps(other, DES_DESCOMP, (k, i)):
        topq = O.neg(tuple(op2.dq))
        j2 = op2.dq[1]
        f = _recoupling(j2, j1, j21, 2 + j2)
        f *= -1.0 if j2 == 0 else 1.0                                      # TensorOp::getTransposeFactorDD, abelian
        if transpose:
            f *= O.commute_parity(iq, kq, topq)
        if other_is_left is False:                                         # loop block is the left child
            f *= O.commute_parity(O.neg(iq), topq, kq)
        tensor_product(pi, O.View(op2, True), O.View(op1, True), other_is_left, c, f * scale)


def build_crecredescomp(pi: ProductInfo, k: int, dq, ints: Integrals, hubbard=False) -> O.Op:
    """CreCreDesComp::build Operators.C:1905-1966."""
    c = allocate(pi, CRE_CRE_DESCOMP, (k,), 0, dq, True)
    if not _carry_over(pi, CRE_CRE_DESCOMP, (k,), dq, c) or hubbard:
        return c
    L, R = pi.left, pi.right
    loop_is_left = L.loop                                                  # assignloopblock BaseOperator.C:298-304
    loopb, otherb = (L, R) if loop_is_left else (R, L)
    if any(op.optype == CRE_DESCOMP for op in loopb.ops):
        for src, holder_is_left in ((loopb, not loop_is_left), (otherb, loop_is_left)):
            for op1 in [op for op in src.ops if op.optype == CRE]:
                _cxcdcomp(pi, holder_is_left, op1, k, c, 1.0)
                _dxcccomp(pi, holder_is_left, op1, k, c, 2.0, ints)        # 2.0: CCcomp_ij = -CCcomp_ji
    return c


def build_ham(pi: ProductInfo, hubbard=False) -> O.Op:
    """Ham::build Operators.C:2322-2395: H_L x 1 + 1 x H_R + the same operator pairs SpinBlock::multiplyH applies to a wavefunction
    (opxop.C:22-143 are the operator forms of :155-285), accumulated with TensorProduct instead of TensorMultiply."""
    c = allocate(pi, HAM, (), 0, (0, 0, 0), False)
    big = O.Big(left=pi.left, right=pi.right, psi_dq=(0, 0, 0), core_energy=0.0, hubbard=hubbard)
    for lop, rop, scale in O.h_terms(big):
        tensor_product(pi, lop, rop, True, c, scale)
    return c
