"""CPU oracle (numpy restatement) of the DMRG sweep hot path of sanshar/Block 1.1.1.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this
module, and only as the checker - never as the thing measured or shipped.

Parity status: PINNED.  Every function below is checked (tests/test_oracle_vs_reference.py) against records
dumped from the real reference compiled here (oracle/Makefile -> oracle/_ref/block_dump, fixtures under
tests/golden/), which itself reproduces the reference's golden energies (dmrg_tests/runtest:12,23,30).

Each function cites the reference file:line it restates.  Conventions: spins are the integers 2S; point group is
abelian (c1, ci, cs, c2, c2v, c2h, d2, d2h: irrep product = XOR, Symmetry.C:78-104,624-627), which covers every
BASELINE config; non-abelian groups are out of scope.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from functools import lru_cache

import numpy as np

# opTypes enum, BaseOperator.h:35-44
HAM, CRE, CRE_CRE, DES_DESCOMP, CRE_DES, CRE_DESCOMP, CRE_CRE_DESCOMP = 0, 1, 2, 3, 4, 5, 6
DES, DES_DES, CRE_CRECOMP, DES_CRE, DES_CRECOMP, CRE_DES_DESCOMP, OVERLAP = 7, 8, 9, 10, 11, 12, 13
HUBBARD_HAM = 1  # hamTypes {QUANTUM_CHEMISTRY, HUBBARD, ...} input.h
NUMERICAL_ZERO = 1e-15  # dmrg.C:85


# ----------------------------------------------------------------------------------------------------------
# angular momentum algebra  (new_anglib.C, couplingCoeffs.h:82-97)
# ----------------------------------------------------------------------------------------------------------
def _fact(n: int) -> float:
    return float(math.factorial(n))


@lru_cache(maxsize=None)
def clebsch(j1: int, m1: int, j2: int, m2: int, j3: int, m3: int) -> float:
    """<j1 m1 j2 m2 | j3 m3>, all arguments doubled.  new_anglib.C:135-206 (Racah's closed form)."""
    if j1 < 0 or j2 < 0 or j3 < 0 or abs(m1) > j1 or abs(m2) > j2 or abs(m3) > j3:
        return 0.0
    if j1 + j2 < j3 or abs(j1 - j2) > j3 or m1 + m2 != m3:
        return 0.0
    if (j1 + m1) % 2 or (j2 + m2) % 2 or (j3 + m3) % 2 or (j1 + j2 + j3) % 2:
        return 0.0
    h = lambda x: x // 2
    pref = math.sqrt((j3 + 1) * _fact(h(j1 + j2 - j3)) * _fact(h(j1 - j2 + j3)) * _fact(h(-j1 + j2 + j3)) / _fact(h(j1 + j2 + j3) + 1))
    pref *= math.sqrt(_fact(h(j1 + m1)) * _fact(h(j1 - m1)) * _fact(h(j2 + m2)) * _fact(h(j2 - m2)) * _fact(h(j3 + m3)) * _fact(h(j3 - m3)))
    s = 0.0
    kmin = max(0, h(j2 - j3 - m1), h(j1 - j3 + m2))
    kmax = min(h(j1 + j2 - j3), h(j1 - m1), h(j2 + m2))
    for k in range(kmin, kmax + 1):
        s += (-1) ** k / (_fact(k) * _fact(h(j1 + j2 - j3) - k) * _fact(h(j1 - m1) - k) * _fact(h(j2 + m2) - k)
                          * _fact(h(j3 - j2 + m1) + k) * _fact(h(j3 - j1 - m2) + k))
    return pref * s


def _tri(a: int, b: int, c: int) -> float:
    return _fact((a + b - c) // 2) * _fact((a - b + c) // 2) * _fact((-a + b + c) // 2) / _fact((a + b + c) // 2 + 1)


@lru_cache(maxsize=None)
def six_j(a: int, b: int, c: int, d: int, e: int, f: int) -> float:
    """Wigner 6j {a b c; d e f}, doubled arguments.  new_anglib.C:76-116."""
    for (x, y, z) in ((a, b, c), (c, d, e), (a, e, f), (b, d, f)):
        if (x + y + z) % 2 or x + y < z or abs(x - y) > z:
            return 0.0
    pref = math.sqrt(_tri(a, b, c) * _tri(c, d, e) * _tri(a, e, f) * _tri(b, d, f))
    h = lambda x: x // 2
    tmin = max(h(a + b + c), h(c + d + e), h(a + e + f), h(b + d + f))
    tmax = min(h(a + b + d + e), h(a + c + d + f), h(b + c + e + f))
    s = 0.0
    for t in range(tmin, tmax + 1):
        s += (-1) ** t * _fact(t + 1) / (_fact(t - h(a + b + c)) * _fact(t - h(c + d + e)) * _fact(t - h(a + e + f)) * _fact(t - h(b + d + f))
                                         * _fact(h(a + b + d + e) - t) * _fact(h(a + c + d + f) - t) * _fact(h(b + c + e + f) - t))
    return pref * s


@lru_cache(maxsize=None)
def nine_j(a, b, c, d, e, f, g, h, i) -> float:
    """Wigner 9j, doubled arguments.  new_anglib.C:19-72 (sum over k of three 6j)."""
    for (x, y, z) in ((a, b, c), (d, e, f), (g, h, i), (a, d, g), (b, e, h), (c, f, i)):
        if x + y < z or abs(x - y) > z:
            return 0.0
    kmin = max(abs(h - d), abs(b - f), abs(a - i))
    kmax = min(h + d, b + f, a + i)
    s = 0.0
    for k in range(kmin, kmax + 1):
        s += (-1) ** k * (k + 1) * six_j(a, b, c, f, i, k) * six_j(d, e, f, b, k, h) * six_j(g, h, i, k, a, d)
    return s


def ninej(ja, jb, jc, jd, je, jf, jg, jh, ji, spin_adapted=True) -> float:
    """ninejCoeffs::operator() couplingCoeffs.C:54-68 == Ninej couplingCoeffs.h:88-97 (table or direct: same value)."""
    if not spin_adapted:
        return 1.0
    return math.sqrt((jg + 1) * (jh + 1) * (jc + 1) * (jf + 1)) * nine_j(ja, jb, jc, jd, je, jf, jg, jh, ji)


def irrep_mul(a: int, b: int) -> int:
    return a ^ b


def qn_allow(q, q1, q2) -> bool:
    """SpinQuantum::allow SpinQuantum.C:99-107: q in q1 (+) q2  (N additive, SU(2) triangle, abelian irrep product)."""
    if q[0] != q1[0] + q2[0] or q[2] != irrep_mul(q1[2], q2[2]):
        return False
    return abs(q1[1] - q2[1]) <= q[1] <= q1[1] + q2[1] and (q1[1] + q2[1] - q[1]) % 2 == 0


def neg(q):
    """SpinQuantum::operator- : (-N, S, -irrep); abelian irreps are self-conjugate (Symmetry.C:374-400)."""
    return (-q[0], q[1], q[2])


def commute_parity(a, b, c) -> float:
    """getCommuteParity BaseOperator.C:20-53 (abelian: spatial factor 1)."""
    parity = -1.0 if (a[0] % 2 and b[0] % 2) else 1.0
    for asz in range(-a[1], a[1] + 1, 2):
        for bsz in range(-b[1], b[1] + 1, 2):
            cleb = clebsch(a[1], asz, b[1], bsz, c[1], c[1])
            if abs(cleb) <= NUMERICAL_ZERO:
                continue
            return parity * cleb / clebsch(b[1], bsz, a[1], asz, c[1], c[1])
    raise ValueError("getCommuteParity: inappropriate operators %s %s %s" % (a, b, c))


def transpose_scaling(opdq, leftq, rightq) -> float:
    """Transposeview::get_scaling BaseOperator.C:56-91 for conjugacy 't' (abelian: spatial factor 1)."""
    ls, rs, cs = leftq[1], rightq[1], opdq[1]
    for lsz in range(-ls, ls + 1, 2):
        for rsz in range(-rs, rs + 1, 2):
            cleb = clebsch(ls, lsz, cs, -cs, rs, rsz)
            if abs(cleb) <= NUMERICAL_ZERO:
                continue
            return (-1.0) ** cs * cleb / clebsch(rs, rsz, cs, cs, ls, lsz)
    raise ValueError("get_scaling: inappropriate arguments")


def transpose_factor_dd(pspin: int) -> float:
    """TensorOp::getTransposeFactorDD tensor_operator.h:164-190 (abelian: spatial factor 1)."""
    return -1.0 if pspin == 0 else 1.0


# ----------------------------------------------------------------------------------------------------------
# data model: mirrors StateInfo / SparseMatrix / SpinBlock only as far as the hot path reads them
# ----------------------------------------------------------------------------------------------------------
@dataclass
class Op:
    """One SparseMatrix (BaseOperator.h:75-243): sector blocks [i][j] row-major."""
    optype: int
    orbs: tuple
    comp: int            # index inside the vector over spin components (op_components.C:172-184)
    dq: tuple            # deltaQuantum[0] = (dN, 2S, irrep)
    fermion: bool
    allowed: np.ndarray  # (nq, nq) bool
    blocks: dict         # (i, j) -> ndarray (d_i, d_j)


@dataclass
class Block:
    q: np.ndarray        # (nq, 3) int: N, 2S, irrep
    dims: np.ndarray     # (nq,)
    ops: list = field(default_factory=list)
    sites: tuple = ()
    loop: bool = False

    def array(self, optype):
        """Operator array of one type in storage order: list of (orbs, [components...])."""
        out, index = [], {}
        for op in self.ops:
            if op.optype != optype:
                continue
            if op.orbs not in index:
                index[op.orbs] = len(out)
                out.append((op.orbs, []))
            out[index[op.orbs]][1].append(op)
        return out

    def get(self, optype, orbs, comp):
        for op in self.ops:
            if op.optype == optype and op.orbs == tuple(orbs) and op.comp == comp:
                return op
        return None


class View:
    """Operator or its Transposeview (BaseOperator.h:208-243) as TensorMultiply sees it."""

    def __init__(self, op: Op, t: bool = False):
        self.op, self.t = op, t
        self.spin = op.dq[1]                      # get_spin: stored spin either way (:232)
        self.fermion = op.fermion

    def allowed(self, i, j):
        return bool(self.op.allowed[j, i] if self.t else self.op.allowed[i, j])

    def mat(self, i, j):
        """The matrix actually multiplied: stored block (j,i) transposed for a view, (i,j) as is otherwise."""
        return self.op.blocks[(j, i)].T if self.t else self.op.blocks[(i, j)]

    def stored(self, i, j):
        return self.op.blocks[(j, i)] if self.t else self.op.blocks[(i, j)]

    def scaling(self, lq, rq):
        return transpose_scaling(self.op.dq, tuple(lq), tuple(rq)) if self.t else 1.0


@dataclass
class Big:
    left: Block
    right: Block
    psi_dq: tuple
    core_energy: float = 0.0
    hubbard: bool = False

    def __post_init__(self):
        L, R = self.left, self.right
        self.allowed = np.zeros((len(L.dims), len(R.dims)), bool)
        for l in range(len(L.dims)):
            for r in range(len(R.dims)):
                self.allowed[l, r] = qn_allow(self.psi_dq, tuple(L.q[l]), tuple(R.q[r]))   # wavefunction.C:31-38
        self.offsets, off = {}, 0
        for l in range(len(L.dims)):
            for r in range(len(R.dims)):
                if self.allowed[l, r]:
                    self.offsets[(l, r)] = off
                    off += int(L.dims[l]) * int(R.dims[r])
        self.size = off

    def unflatten(self, flat):
        """Wavefunction::CollectFrom wavefunction.C:188-203."""
        L, R = self.left, self.right
        return {k: np.array(flat[o:o + int(L.dims[k[0]]) * int(R.dims[k[1]])], dtype=np.float64).reshape(int(L.dims[k[0]]), int(R.dims[k[1]]))
                for k, o in self.offsets.items()}

    def flatten(self, w):
        """Wavefunction::FlattenInto wavefunction.C:167-186."""
        out = np.zeros(self.size)
        for k, o in self.offsets.items():
            out[o:o + w[k].size] = w[k].ravel()
        return out

    def zeros(self):
        return self.unflatten(np.zeros(self.size))


# ----------------------------------------------------------------------------------------------------------
# sigma = H psi
# ----------------------------------------------------------------------------------------------------------
def tensor_multiply(big: Big, lop: View, rop: View, c, v, opq_spin: int, scale: float, flops=None):
    """operatorfunctions::TensorMultiply(ablock,a,b,cblock,c,v,opQ,scale), operatorfunctions.C:485-537,
    written with leftOp/rightOp already resolved (:496-501)."""
    L, R = big.left, big.right
    S = big.psi_dq[1]
    nl, nr = len(L.dims), len(R.dims)
    for lQ in range(nl):
        for rQp in range(nr):
            for lQp in range(nl):
                if not (lop.allowed(lQ, lQp) and big.allowed[lQp, rQp]):
                    continue
                m = lop.scaling(L.q[lQ], L.q[lQp]) * (lop.mat(lQ, lQp) @ c[(lQp, rQp)])            # :512-516
                if flops is not None:
                    flops[0] += 2.0 * L.dims[lQ] * L.dims[lQp] * R.dims[rQp]
                for rQ in range(nr):
                    if not (big.allowed[lQ, rQ] and rop.allowed(rQ, rQp)):
                        continue
                    f = scale * ninej(L.q[lQp][1], R.q[rQp][1], S, lop.spin, rop.spin, opq_spin, L.q[lQ][1], R.q[rQ][1], S)   # :522-524
                    if rop.fermion and L.q[lQp][0] % 2:                                              # :528
                        f = -f
                    f *= rop.scaling(R.q[rQ], R.q[rQp])                                              # :529
                    v[(lQ, rQ)] += f * (m @ rop.mat(rQ, rQp).T)                                      # :530
                    if flops is not None:
                        flops[0] += 2.0 * L.dims[lQ] * R.dims[rQp] * R.dims[rQ]


def h_terms(big: Big):
    """Term list of SpinBlock::multiplyH spinblock.C:722-789 with the functors of opxop.C:155-285 expanded
    (energy sweep: implicitTranspose, no DES/DES_CRE/..COMP-transposed arrays).  Yields (leftView, rightView, scale)."""
    L, R = big.left, big.right
    hq = (0, 0, 0)
    terms = []

    def pair(other_is_left, op_other, t_other, op_loop, t_loop, scale):
        a, b = View(op_other, t_other), View(op_loop, t_loop)
        terms.append((a, b, scale) if other_is_left else (b, a, scale))

    ovl_l, ovl_r = L.get(OVERLAP, (), 0), R.get(OVERLAP, (), 0)
    ham_l, ham_r = L.get(HAM, (), 0), R.get(HAM, (), 0)
    if abs(big.core_energy) > 1e-20:                                   # :735-740 (TINY = 1e-20, global.h)
        terms.append((View(ovl_l), View(ovl_r), big.core_energy))
    terms.append((View(ham_l), View(ovl_r), 1.0))                     # :742-744
    terms.append((View(ovl_l), View(ham_r), 1.0))                     # :745-747

    # c x ccd_comp, both directions  (:757-763 -> opxop.C:232-285, else-branch)
    for other, loopb, other_is_left in ((L, R, True), (R, L, False)):
        for orbs, comps in loopb.array(CRE):
            for k, op1 in enumerate(comps):
                op2 = other.get(CRE_CRE_DESCOMP, orbs, k)
                if op2 is None:
                    break                                              # has_local_index false -> return (:240)
                par = commute_parity(neg(op1.dq), op2.dq, hq) if not other_is_left else 1.0     # :273-274
                pair(other_is_left, op2, False, op1, True, par)                                      # :276
                par = commute_parity(op1.dq, neg(op2.dq), hq) if other_is_left else 1.0         # :278-279
                pair(other_is_left, op2, True, op1, False, par)                                      # :282
    if not big.hubbard:                                                # :771
        loopb, other = (L, R) if L.loop else (R, L)
        other_is_left = other is L
        for orbs, comps in loopb.array(CRE_DES):                       # cdxcdcomp opxop.C:155-185
            for k, op1 in enumerate(comps):
                op2 = other.get(CRE_DESCOMP, orbs, k)
                if op2 is None:
                    break
                pair(other_is_left, op2, False, op1, False, 1.0)
                if orbs[0] != orbs[1]:
                    pair(other_is_left, op2, True, op1, True, 1.0)
        for orbs, comps in loopb.array(CRE_CRE):                       # ddxcccomp opxop.C:187-228
            for k, op1 in enumerate(comps):
                op2 = other.get(DES_DESCOMP, orbs, k)
                if op2 is None:
                    break
                factor = 1.0 if orbs[0] == orbs[1] else 2.0
                par = commute_parity(op1.dq, op2.dq, hq) if other_is_left else 1.0
                pair(other_is_left, op2, False, op1, False, factor * par)
                par *= transpose_factor_dd(op1.dq[1]) * transpose_factor_dd(op2.dq[1])
                pair(other_is_left, op2, True, op1, True, factor * par)
    return terms


def multiply_h(big: Big, c, flops=None):
    """sigma = H c as a fresh sector dict (the reference accumulates into a cleared v, linear.C:239-253)."""
    v = big.zeros()
    for lop, rop, scale in h_terms(big):
        tensor_multiply(big, lop, rop, c, v, 0, scale, flops)
    return v


def diagonal_h(big: Big):
    """SpinBlock::diagonalH spinblock.C:855-899 -> TensorTrace/TensorProduct diagonal forms operatorfunctions.C:653-762
    and the *_d functors opxop.C:295-365.  Returned in flat psi order (big is never 'collected')."""
    L, R = big.left, big.right
    S = big.psi_dq[1]
    e = np.zeros(big.size)

    def trace_left(a: View, scale):       # TensorTrace conjC == 'n' (:672-681)
        for (l, r), off in big.offsets.items():
            if not a.allowed(l, l):
                continue
            f = scale * ninej(L.q[l][1], R.q[r][1], S, a.spin, 0, 0, L.q[l][1], R.q[r][1], S)
            d = np.diag(a.stored(l, l))
            e[off:off + d.size * R.dims[r]] += np.repeat(f * d, R.dims[r])

    def trace_right(a: View, scale):      # TensorTrace conjC == 't' (:682-692)
        for (l, r), off in big.offsets.items():
            if not a.allowed(r, r):
                continue
            f = scale * ninej(L.q[l][1], R.q[r][1], S, 0, a.spin, 0, L.q[l][1], R.q[r][1], S)
            if a.fermion and L.q[l][0] % 2:
                f = -f
            d = np.diag(a.stored(r, r))
            e[off:off + d.size * L.dims[l]] += np.tile(f * d, L.dims[l])

    def product(a: View, b: View, scale):  # TensorProduct(..., DiagonalMatrix) with a on the left (:727-743 / :744-759)
        for (l, r), off in big.offsets.items():
            if not (a.allowed(l, l) and b.allowed(r, r)):
                continue
            f = scale * ninej(L.q[l][1], R.q[r][1], S, a.spin, b.spin, 0, L.q[l][1], R.q[r][1], S)
            if b.fermion and L.q[l][0] % 2:
                f = -f
            e[off:off + L.dims[l] * R.dims[r]] += f * np.outer(np.diag(a.stored(l, l)), np.diag(b.stored(r, r))).ravel()

    trace_left(View(L.get(HAM, (), 0)), 1.0)
    trace_right(View(R.get(HAM, (), 0)), 1.0)
    e += big.core_energy

    def pair(other_is_left, op_other, t_other, op_loop, t_loop, scale):
        a, b = View(op_other, t_other), View(op_loop, t_loop)
        product(a, b, scale) if other_is_left else product(b, a, scale)

    for other, loopb, other_is_left in ((L, R, True), (R, L, False)):      # cxcddcomp_d opxop.C:341-365
        for orbs, comps in loopb.array(CRE):
            for k, op1 in enumerate(comps):
                op2 = other.get(CRE_CRE_DESCOMP, orbs, k)
                if op2 is None:
                    break
                pair(other_is_left, op2, False, op1, True, 1.0)
                pair(other_is_left, op2, True, op1, False, 1.0)
    if not big.hubbard:
        loopb, other = (L, R) if L.loop else (R, L)
        other_is_left = other is L
        for orbs, comps in loopb.array(CRE_DES):                           # cdxcdcomp_d opxop.C:295-312
            for k, op1 in enumerate(comps):
                op2 = other.get(CRE_DESCOMP, orbs, k)
                if op2 is None:
                    break
                pair(other_is_left, op2, False, op1, False, 1.0)
                if orbs[0] != orbs[1]:
                    pair(other_is_left, op2, True, op1, True, 1.0)
        for orbs, comps in loopb.array(CRE_CRE):                           # ddxcccomp_d opxop.C:314-339
            for k, op1 in enumerate(comps):
                op2 = other.get(DES_DESCOMP, orbs, k)
                if op2 is None:
                    break
                factor = 1.0 if orbs[0] == orbs[1] else 2.0
                pair(other_is_left, op2, False, op1, False, factor)
                pair(other_is_left, op2, True, op1, True, factor)
    return e


# ----------------------------------------------------------------------------------------------------------
# Davidson  (linear.C:27-60, 179-385) on flat vectors
# ----------------------------------------------------------------------------------------------------------
def precondition(op, e, diag):
    """Linear::precondition linear.C:27-42 (levelshift 0)."""
    den = e - diag
    mask = np.abs(den) > 1e-12
    out = op.copy()
    out[mask] /= den[mask]
    return out


def olsen_precondition(r, c0, e, diag):
    """Linear::olsenPrecondition linear.C:50-60."""
    c0p = precondition(c0, e, diag)
    r = r - (np.dot(c0p, r) / np.dot(c0, c0p)) * c0
    return precondition(r, e, diag)


def block_davidson(hmul, guesses, diag, tol, defl_min=2, defl_max=20, max_iter=10000, lower=()):
    """Linear::block_davidson linear.C:179-385.  `lower` = lowerStates of a state-specific solve (currentRoot >= 0), already
    orthogonalised among themselves by the caller (solver.C:79-86); empty for the state-averaged form.
    hmul: flat -> flat.  Returns (eigenvalues[nroots], vectors, number of H applications)."""
    b = [np.array(g, dtype=np.float64) for g in guesses]
    lower = [np.asarray(l, dtype=np.float64) for l in lower]
    nroots = len(b)

    def project(r):                                               # :203-206, :311-317, :369-375
        for l in lower:
            r = r - (np.dot(r, l) / np.dot(l, l)) * l
        return r
    for i in range(nroots):                                       # :190-198
        for j in range(i):
            b[i] = b[i] - np.dot(b[j], b[i]) * b[j]
        b[i] = b[i] / math.sqrt(np.dot(b[i], b[i]))
    if lower:                                                     # :201-208: only b[0]
        b[0] = project(b[0])
        b[0] = b[0] / math.sqrt(np.dot(b[0], b[0]))
    sigma, converged, nmult = [], 0, 0
    for _ in range(max_iter):
        for i in range(len(sigma), len(b)):                       # :234-257
            sigma.append(hmul(b[i])); nmult += 1
        n = len(b)
        hs = np.zeros((n, n))
        for i in range(n):
            for j in range(i + 1):
                hs[i, j] = hs[j, i] = np.dot(b[i], sigma[j])      # :266-270
        theta, alpha = np.linalg.eigh(hs)                         # :273 (dsyev, ascending)
        B, Sg = np.array(b), np.array(sigma)
        b = list(alpha.T @ B)                                     # :279-294 Ritz rotation of b and sigma
        sigma = list(alpha.T @ Sg)
        for i in range(converged):                                # :298-307
            r = sigma[i] - theta[i] * b[i]
            if np.dot(r, r) > tol:
                converged = i
        r = project(sigma[converged] - theta[converged] * b[converged])
        rnorm = np.dot(r, r)                                      # :321-323, of the projected residual
        r = olsen_precondition(r, b[converged], theta[converged], diag)   # :331
        if rnorm < tol:                                           # :335
            converged += 1
            if converged == nroots:
                return theta[:nroots].copy(), b[:nroots], nmult
            continue
        if len(b) >= defl_max:                                    # :352-357
            b, sigma = b[:defl_min], sigma[:defl_min]
        for j in range(len(b)):                                   # :358-366
            r = r / math.sqrt(np.dot(r, r))
            r = r - np.dot(r, b[j]) * b[j]
        r = project(r)
        r = r / math.sqrt(np.dot(r, r))
        b.append(r)
    raise RuntimeError("davidson did not converge")


# ----------------------------------------------------------------------------------------------------------
# renormalisation  (density.C:27-90, rotationmat.C:149-346, renormalise.C:135-166, BaseOperator.C:341-363)
# ----------------------------------------------------------------------------------------------------------
def make_density(big: Big, waves, weights):
    """DensityMatrix::makedensitymatrix density.C:27-33,84-90 -> MultiplyProduct operatorfunctions.C:630-650 (noise 0)."""
    L = big.left
    rho = [np.zeros((int(d), int(d))) for d in L.dims]
    for w, wt in zip(waves, weights):
        if abs(wt) < 1e-20:
            continue
        for (l, r), _ in big.offsets.items():
            rho[l] += wt * (w[(l, r)] @ w[(l, r)].T)
    return rho


class WaveLayout:
    """Sector layout of a Wavefunction with target quantum dq on `big` (Wavefunction::initialise wavefunction.C:18-56):
    the noise wavefunctions O.psi live in sectors shifted by the operator's quantum numbers."""

    def __init__(self, big: Big, dq):
        L, R = big.left, big.right
        self.big, self.dq = big, tuple(int(x) for x in dq)
        self.offsets, off = {}, 0
        for l in range(len(L.dims)):
            for r in range(len(R.dims)):
                if qn_allow(self.dq, tuple(L.q[l]), tuple(R.q[r])):
                    self.offsets[(l, r)] = off
                    off += int(L.dims[l]) * int(R.dims[r])
        self.size = off

    def allowed(self, l, r):
        return (l, r) in self.offsets

    def zeros(self):
        L, R = self.big.left, self.big.right
        return {k: np.zeros((int(L.dims[k[0]]), int(R.dims[k[1]]))) for k in self.offsets}

    def flatten(self, w):
        out = np.zeros(self.size)
        for k, o in self.offsets.items():
            out[o:o + w[k].size] = w[k].ravel()
        return out

    def unflatten(self, flat):
        L, R = self.big.left, self.big.right
        return {k: np.array(flat[o:o + int(L.dims[k[0]]) * int(R.dims[k[1]])], dtype=np.float64).reshape(int(L.dims[k[0]]), int(R.dims[k[1]]))
                for k, o in self.offsets.items()}


def tensor_multiply_one(big: Big, a: View, a_is_left: bool, c, c_layout: WaveLayout, v, v_layout: WaveLayout, scale: float):
    """One-operator operatorfunctions::TensorMultiply(ablock, a, cblock, c, v, dQ, scale), operatorfunctions.C:331-404:
    v += scale (a x 1) c (a on the left child, :343-372) or scale (1 x a) c (right child, :374-402)."""
    L, R = big.left, big.right
    Sc, Sv = c_layout.dq[1], v_layout.dq[1]
    nl, nr = len(L.dims), len(R.dims)
    if a_is_left:
        for lQ in range(nl):
            for lQp in range(nl):
                if not a.allowed(lQ, lQp):
                    continue
                for rQ in range(nr):
                    if c_layout.allowed(lQp, rQ) and v_layout.allowed(lQ, rQ):
                        fac = scale * ninej(L.q[lQp][1], R.q[rQ][1], Sc, a.spin, 0, a.spin, L.q[lQ][1], R.q[rQ][1], Sv)     # :357-359
                        fac *= a.scaling(L.q[lQ], L.q[lQp])                                                                  # :363
                        v[(lQ, rQ)] += fac * (a.mat(lQ, lQp) @ c[(lQp, rQ)])                                               # :364
    else:
        for rQ in range(nr):
            for rQp in range(nr):
                if not a.allowed(rQ, rQp):
                    continue
                for lQp in range(nl):
                    if v_layout.allowed(lQp, rQ) and c_layout.allowed(lQp, rQp):
                        fac = scale * ninej(L.q[lQp][1], R.q[rQp][1], Sc, 0, a.spin, a.spin, L.q[lQp][1], R.q[rQ][1], Sv)   # :386-388
                        fac *= a.scaling(R.q[rQ], R.q[rQp])                                                                  # :392
                        if a.fermion and L.q[lQp][0] % 2:                                                                    # :393
                            fac = -fac
                        v[(lQp, rQ)] += fac * (c[(lQp, rQp)] @ a.mat(rQ, rQp).T)                                           # :395


def noise_operator_types(left: Block):
    """Operator arrays add_onedot_noise loops over, density.C:360-379."""
    types = [CRE] if left.array(CRE) else []
    if left.array(CRE_CRE):
        types += [CRE_CRE, CRE_DES]
    elif left.array(DES_DESCOMP):
        types += [DES_DESCOMP, CRE_DESCOMP]
    return types


def add_onedot_noise(big: Big, rho, wave, noise: float):
    """DensityMatrix::add_onedot_noise density.C:332-399 with the functor onedot_noise_f :181-258:
    rho += noise / tr(rho_n) * rho_n,  rho_n = sum_O (O psi)(O psi)^T / |O psi|^2 over the left block's CRE, CRE_CRE, CRE_DES
    (or DES_DESCOMP, CRE_DESCOMP) operators and their transposes, each into the +-dQ shifted target sector.
    Reference quirk kept: for HUBBARD the accumulated rho_n is never added (:358-392)."""
    if big.hubbard:
        return
    L = big.left
    wl = WaveLayout(big, big.psi_dq)
    dmn = [np.zeros((int(d), int(d))) for d in L.dims]
    wq = big.psi_dq
    for optype in noise_operator_types(L):
        for orbs, comps in L.array(optype):
            for op in comps:
                oq = op.dq
                irrep = irrep_mul(wq[2], oq[2])
                for spin in range(abs(wq[1] - oq[1]), wq[1] + oq[1] + 1, 2):                      # spinvec = wQ.s + oQ.s (:208)
                    for sign, view in ((+1, View(op, False)), (-1, View(op, True))):              # :214-224 and :226-236
                        vl = WaveLayout(big, (wq[0] + sign * oq[0], spin, irrep))
                        opx = vl.zeros()
                        tensor_multiply_one(big, view, True, wave, wl, opx, vl, 1.0)
                        norm = sum(float(np.vdot(m, m)) for m in opx.values())
                        if abs(norm) > NUMERICAL_ZERO:
                            inv = 1.0 / math.sqrt(norm)
                            for (l, r), m in opx.items():
                                dmn[l] += (inv * m) @ (inv * m).T                                 # MultiplyProduct(opxwave, Transpose(opxwave), dm, 1.0)
    norm = sum(float(np.trace(m)) for m in dmn)
    if norm > 1.0:                                                                                # :388-389
        for q in range(len(rho)):
            rho[q] += (noise / norm) * dmn[q]


def make_density_with_noise(big: Big, waves, weights, noise: float):
    """DensityMatrix::makedensitymatrix density.C:27-82 (additional_noise = 0)."""
    rho = make_density(big, waves, weights)
    if noise > NUMERICAL_ZERO:
        for w in waves:
            add_onedot_noise(big, rho, w, noise / len(waves))                                     # :57-60
    return rho


def diagonalise_dm(rho):
    """diagonalise_dm rotationmat.C:258-279: per-sector dsyev ascending, eigenvalues < 1e-14 -> 0."""
    evals, evecs = [], []
    for m in rho:
        if m.shape[0] == 0:
            evals.append(np.zeros(0)); evecs.append(np.zeros((0, 0))); continue
        w, v = np.linalg.eigh(m)
        w = np.where(w < 1e-14, 0.0, w)
        evals.append(w); evecs.append(v)
    return evals, evecs


def select_states(evals, keep):
    """sort_weights rotationmat.C:313-346 + assign_matrix_by_dm :149-211 (keptqstates = 0).
    Returns per-sector list of kept eigenvector indices IN SELECTION ORDER, and the discarded weight."""
    entries = []
    for q, w in enumerate(evals):
        for s in range(len(w)):
            entries.append((w[s], len(entries), q, s))
    # multimap reverse iteration: descending key; equal keys come out in reverse insertion order
    entries.sort(key=lambda t: (t[0], t[1]), reverse=True)
    total = min(len(entries), keep)
    kept = [[] for _ in evals]
    norm_kept = 0.0
    for i in range(total):
        w, _, q, s = entries[i]
        if w > 1e-13:
            kept[q].append(s)
            norm_kept += w
    norm = sum(float(np.sum(w)) for w in evals)
    return kept, norm - norm_kept


def rotation_matrices(evecs, kept):
    return [evecs[q][:, kept[q]] if len(kept[q]) else np.zeros((evecs[q].shape[0], 0)) for q in range(len(evecs))]


def rotate_op(op: Op, rot):
    """SparseMatrix::renormalise_transform BaseOperator.C:341-363 -> MatrixRotate MatrixBLAS.C:553-572:
    O'[a,b] = U_Q(a)^T O[Q(a),Q(b)] U_Q(b) over the sectors that kept >= 1 state (save_load_block.C:270-283)."""
    keepq = [q for q in range(len(rot)) if rot[q].shape[1] > 0]
    n = len(keepq)
    allowed = np.zeros((n, n), bool)
    blocks = {}
    for a, Q in enumerate(keepq):
        for b, Qp in enumerate(keepq):
            if op.allowed[Q, Qp]:
                allowed[a, b] = True
                blocks[(a, b)] = rot[Q].T @ op.blocks[(Q, Qp)] @ rot[Qp]
    return Op(op.optype, op.orbs, op.comp, op.dq, op.fermion, allowed, blocks)


def sigma_flops(big: Big) -> float:
    """ALGORITHMIC flops of one multiplyH: the dgemm flops the reference issues."""
    fl = [0.0]
    L, R = big.left, big.right
    for lop, rop, _ in h_terms(big):
        for lQ in range(len(L.dims)):
            for rQp in range(len(R.dims)):
                for lQp in range(len(L.dims)):
                    if not (lop.allowed(lQ, lQp) and big.allowed[lQp, rQp]):
                        continue
                    fl[0] += 2.0 * L.dims[lQ] * L.dims[lQp] * R.dims[rQp]
                    for rQ in range(len(R.dims)):
                        if big.allowed[lQ, rQ] and rop.allowed(rQ, rQp):
                            fl[0] += 2.0 * L.dims[lQ] * R.dims[rQp] * R.dims[rQ]
    return fl[0]
