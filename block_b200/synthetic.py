"""Synthetic `big` blocks of the shape BASELINE.json names ("synthetic random-integral FCIDUMP, 40 orbitals /
40 electrons, M=4000"): the sector tables and operator arrays a two-dot block iteration of the reference hands to
SpinBlock::multiplyH / RenormaliseFrom, without running the (out-of-scope) block construction.

What is modelled after the reference (file:line under the reference root):
  * a renormalised block of n sites keeps M states spread over (N, 2S) sectors (StateInfo.h:113-147); the
    enlarged block is (renormalised block) x (one site: (0,0), (1,1), (2,0)), quanta collected and sorted by
    (N, 2S, irrep) (StateInfo.C TensorProduct + CollectQuanta, SpinQuantum operator<);
  * the operator arrays a block carries in the energy sweep (set_spinblock_components.C:490-560): HAM, OVERLAP,
    CRE_i for its own sites, CRE_CRE_DESCOMP_i for the other block's sites; the loop block carries CRE_DES_ij /
    CRE_CRE_ij (i >= j own sites, spin components S = 0, 1: op_components.C:172-184), the other block the matching
    CRE_DESCOMP_ij / DES_DESCOMP_ij;
  * an operator block (i, j) is allocated iff q_i is in deltaQuantum x q_j (SparseMatrix::allocate,
    BaseOperator.C:123-145); C1 symmetry (ORBSYM all 1), so irreps are 0 throughout.
Operator VALUES are counter-based random numbers filled on the device (b2d_fill_op_random): sigma throughput does
not depend on them, and HAM is made self-adjoint so that Davidson runs on a symmetric matrix.
"""
from __future__ import annotations

import math

import numpy as np

from .hotpath import BlockSpec, OperatorSpec, SpinBlock

HAM, CRE, CRE_CRE, DES_DESCOMP, CRE_DES, CRE_DESCOMP, CRE_CRE_DESCOMP, OVERLAP = 0, 1, 2, 3, 4, 5, 6, 13


def csf_count(n, N, twoS):
    """Number of spin-adapted configurations of N electrons in n orbitals with spin S (Weyl's formula): the cap on
    a sector's size."""
    if N < 0 or N > 2 * n or twoS < 0 or (N - twoS) % 2 or twoS > min(N, 2 * n - N):
        return 0
    a = (N - twoS) // 2
    b = (N + twoS) // 2 + 1
    return (twoS + 1) * math.comb(n + 1, a) * math.comb(n + 1, b) // (n + 1)


def renormalised_sectors(nsites, nelec_mean, M, sigma_n=2.2, sigma_s=1.6):
    """{(N, 2S): states} of a renormalised block: Gaussian in N around the mean filling, (2S+1) exp(-S(S+1)/2s^2) in
    spin, scaled so that the sectors add up to M (every populated sector keeps >= 1 state: the long tail of tiny
    sectors is what makes the contraction ragged)."""
    w = {}
    for N in range(0, 2 * nsites + 1):
        for twoS in range(N % 2, min(N, 2 * nsites - N) + 1, 2):
            cap = csf_count(nsites, N, twoS)
            if cap == 0:
                continue
            S = 0.5 * twoS
            w[(N, twoS)] = (math.exp(-0.5 * ((N - nelec_mean) / sigma_n) ** 2) * (twoS + 1) * math.exp(-0.5 * S * (S + 1) / sigma_s ** 2), cap)
    tot = sum(v[0] for v in w.values())
    scale = M / tot
    for _ in range(40):   # fixed point: sectors that round to 0 drop out, capped sectors give their share back
        dims = {k: min(cap, int(round(scale * x))) for k, (x, cap) in w.items()}
        got = sum(dims.values())
        if got == M:
            break
        scale *= M / max(got, 1)
    dims = {k: d for k, d in dims.items() if d > 0}
    # exact total: adjust the largest sector
    kmax = max(dims, key=dims.get)
    dims[kmax] += M - sum(dims.values())
    return dims


def add_dot(sectors):
    """(block) x (one site), quanta collected: dims of the enlarged block, sorted by (N, 2S)."""
    out = {}
    for (N, s), d in sectors.items():
        for (dn, ds) in ((0, 0), (1, 1), (2, 0)):
            for s2 in ([s] if ds == 0 else [s - 1, s + 1]):
                if s2 < 0:
                    continue
                out[(N + dn, s2)] = out.get((N + dn, s2), 0) + d
    return dict(sorted(out.items()))


def allowed_mask(q, dq):
    """SparseMatrix::allocate (BaseOperator.C:123-145): block (i, j) exists iff q_i is in dq x q_j."""
    N, S = q[:, 0], q[:, 1]
    ok_n = N[:, None] == N[None, :] + dq[0]
    ok_s = (np.abs(S[None, :] - dq[1]) <= S[:, None]) & (S[:, None] <= S[None, :] + dq[1]) & ((S[:, None] + S[None, :] + dq[1]) % 2 == 0)
    return (ok_n & ok_s).astype(np.uint8)


def _op(q, optype, orbs, comp, dq, fermion):
    return OperatorSpec(optype=optype, orbs=tuple(orbs), comp=comp, dq=tuple(dq), fermion=fermion, allowed=allowed_mask(q, dq), data=None)


def make_block(sectors, sites, other_sites, loop):
    q = np.array([[N, s, 0] for (N, s) in sectors], dtype=np.int32)
    dims = np.array(list(sectors.values()), dtype=np.int32)
    blk = BlockSpec(q=q, dims=dims, sites=tuple(sites), loop=loop)
    blk.ops.append(_op(q, HAM, (), 0, (0, 0, 0), False))
    ovl = _op(q, OVERLAP, (), 0, (0, 0, 0), False)       # the identity: diagonal blocks only
    ovl.data = np.concatenate([np.eye(int(d)).ravel() for d in dims])
    blk.ops.append(ovl)
    for i in sites:
        blk.ops.append(_op(q, CRE, (i,), 0, (1, 1, 0), True))
    for i in other_sites:
        blk.ops.append(_op(q, CRE_CRE_DESCOMP, (i,), 0, (1, 1, 0), True))
    pair_sites = sites if loop else other_sites
    for a, i in enumerate(pair_sites):
        for j in pair_sites[:a + 1]:
            for comp, s in ((0, 0), (1, 2)):
                if loop:
                    blk.ops.append(_op(q, CRE_DES, (i, j), comp, (0, s, 0), False))
                    blk.ops.append(_op(q, CRE_CRE, (i, j), comp, (2, s, 0), False))
                else:
                    blk.ops.append(_op(q, CRE_DESCOMP, (i, j), comp, (0, s, 0), False))
                    blk.ops.append(_op(q, DES_DESCOMP, (i, j), comp, (-2, s, 0), False))
    return blk


def make_big_block(norbs=40, nelec=40, M=4000, left_sites=None, device=0, rank=0, nranks=1, options=None, seed=20260, fill=True,
                   sigma_n=2.2, sigma_s=1.6):
    """The big block of the block iteration with `left_sites` orbitals on the left (default: the middle of the
    chain), both children enlarged blocks of renormalised M-state blocks.  The left child is the loop block.
    device = -1 gives a planning-only context (flop / memory accounting on a CPU)."""
    nl = norbs // 2 if left_sites is None else left_sites
    nr = norbs - nl
    filling = nelec / norbs
    L = add_dot(renormalised_sectors(nl - 1, filling * (nl - 1), M, sigma_n, sigma_s))
    R = add_dot(renormalised_sectors(nr - 1, filling * (nr - 1), M, sigma_n, sigma_s))
    # only sectors that can pair up to the target (nelec, S = 0) survive in a converged calculation
    L = {k: d for k, d in L.items() if (nelec - k[0], k[1]) in R}
    R = {k: d for k, d in R.items() if (nelec - k[0], k[1]) in L}
    lsites, rsites = list(range(nl)), list(range(nl, norbs))
    left = make_block(L, lsites, rsites, loop=True)
    right = make_block(R, rsites, lsites, loop=False)
    sb = SpinBlock(left, right, (nelec, 0, 0), core_energy=0.0, hubbard=False, norbs=norbs, device=device, rank=rank, nranks=nranks, options=options)
    sb.fill_seed = seed
    if fill and device >= 0:
        fill_random(sb, seed, rank, nranks)
    return sb


def fill_params(dims, side, k, optype, seed=20260):
    """(stream seed, amplitude) of operator k of a side: amplitudes ~ 1/sqrt(dimension) keep sigma O(1)."""
    amp = 1.0 / math.sqrt(float(np.sum(dims)))
    return seed * 1000003 + side * 500009 + k, amp * (4.0 if optype == HAM else 1.0)


def fill_random(sb: SpinBlock, seed, rank=0, nranks=1):
    """Counter-based random operator values on the device (operators a rank does not hold are skipped by the library);
    the stream is a pure function of (seed, element index), so every rank - and oracle/ref_bench - sees the same values."""
    for side, blk in enumerate((sb.left, sb.right)):
        for k, op in enumerate(blk.ops):
            if op.data is not None:
                continue
            s, amp = fill_params(blk.dims, side, k, op.optype, seed)
            sb.fill_op_random(side, sb.op_ids[side][k], s, amplitude=amp, symmetric=op.optype == HAM)


# ---- guess-wavefunction transform of the same shape ----------------------------------------------------------------
DOT = ((0, 0), (1, 1), (2, 0))     # sectors of one spin-adapted site: (N, 2S), one state each


def product_tables(sectors, dims):
    """StateInfo tables of (block) x (one site) as b2d_stateinfo wants them: un-collected pieces in (block sector outer, dot sector
    inner, coupled spin increasing) order, collected quanta sorted by (N, 2S); C1 symmetry."""
    unc_q, unc_dims, lmap, rmap = [], [], [], []
    for a, (N, s) in enumerate(sectors):
        for b, (dn, ds) in enumerate(DOT):
            for s2 in ([s] if ds == 0 else [s - 1, s + 1]):
                if s2 < 0:
                    continue
                unc_q.append((N + dn, s2, 0)); unc_dims.append(int(dims[a])); lmap.append(a); rmap.append(b)
    keys = sorted(set(unc_q))
    index = {k: i for i, k in enumerate(keys)}
    pieces = [[] for _ in keys]
    for u, k in enumerate(unc_q):
        pieces[index[k]].append(u)
    begin = np.zeros(len(keys) + 1, np.int32)
    begin[1:] = np.cumsum([len(p) for p in pieces])
    return {"q": np.array(keys, np.int32), "dims": np.array([sum(unc_dims[u] for u in p) for p in pieces], np.int32),
            "unc.q": np.array(unc_q, np.int32), "unc.dims": np.array(unc_dims, np.int32), "unc.lmap": np.array(lmap, np.int32),
            "unc.rmap": np.array(rmap, np.int32), "old_to_new": np.array([u for p in pieces for u in p], np.int32), "old_to_new_begin": begin}


def make_guess_case(norbs=40, nelec=40, M=4000, left_sites=None, seed=20260, sigma_n=2.2, sigma_s=1.6):
    """Inputs of b2d_guess_plan / b2d_guess_transform for the block iteration make_big_block describes: previous wavefunction
    [S (x) d1][E_old (x) d2] -> trial vector [S' (x) d2][d3 (x) E''] (guess_wavefunction.C:524-636).  Returns (dq, tables, old_allowed,
    lrot_cols, rrot_cols, old_wave, left_rot, right_rot); values are seeded random numbers (throughput does not depend on them)."""
    nl = norbs // 2 if left_sites is None else left_sites
    nr = norbs - nl
    f = nelec / norbs
    rs = lambda n: renormalised_sectors(n, f * n, M, sigma_n, sigma_s)
    oldleft = add_dot(rs(nl - 2))                       # S (x) d1, collected: the basis the left rotation matrix truncates
    sys = {k: min(d, oldleft[k]) for k, d in rs(nl - 1).items() if k in oldleft}      # S': M states
    right = add_dot(rs(nr - 1))                         # d3 (x) E'', collected: the basis the right rotation matrix truncated
    env = {k: min(d, right[k]) for k, d in rs(nr).items() if k in right}             # E_old: M states
    q3 = lambda d: np.array([[N, s, 0] for (N, s) in d], np.int32)
    dims = lambda d: np.array(list(d.values()), np.int32)
    okeys, rkeys = list(oldleft), list(right)
    tables = {
        "sys": {"q": q3(sys), "dims": dims(sys), "new_quanta_map": np.array([okeys.index(k) for k in sys], np.int32)},
        "dot": {"q": np.array([[n, s, 0] for (n, s) in DOT], np.int32), "dims": np.ones(3, np.int32)},
        "left": product_tables(list(sys), dims(sys)),
        "right": {"q": q3(right), "dims": dims(right)},
        "oldleft": {"q": q3(oldleft), "dims": dims(oldleft)},
        "oldright": product_tables(list(env), dims(env)),
        "env": {"q": q3(env), "dims": dims(env), "new_quanta_map": np.array([rkeys.index(k) for k in env], np.int32)},
    }
    dq = (nelec, 0, 0)
    ol, orr = tables["oldleft"], tables["oldright"]
    old_allowed = (allowed_mask_pair(ol["q"], orr["q"], dq)).astype(np.uint8)
    lrot_cols = np.zeros(len(okeys), np.int32)
    for k, m in sys.items():
        lrot_cols[okeys.index(k)] = m
    rrot_cols = np.zeros(len(rkeys), np.int32)
    for k, m in env.items():
        rrot_cols[rkeys.index(k)] = m
    rng = np.random.default_rng(seed)
    n_old = int((old_allowed * np.outer(ol["dims"], orr["dims"])).sum())
    old_wave = rng.standard_normal(n_old) / math.sqrt(max(n_old, 1))
    left_rot = rng.standard_normal(int((ol["dims"] * lrot_cols).sum())) * 0.02
    right_rot = rng.standard_normal(int((tables["right"]["dims"] * rrot_cols).sum())) * 0.02
    return dq, tables, old_allowed, lrot_cols, rrot_cols, old_wave, left_rot, right_rot


def allowed_mask_pair(ql, qr, dq):
    """Wavefunction::AllowQuantaFor (wavefunction.C:393-415): block (i, j) exists iff dq is in q_i x q_j (C1)."""
    Nl, Sl, Nr, Sr = ql[:, 0][:, None], ql[:, 1][:, None], qr[:, 0][None, :], qr[:, 1][None, :]
    return (Nl + Nr == dq[0]) & (np.abs(Sl - Sr) <= dq[1]) & (dq[1] <= Sl + Sr) & ((Sl + Sr + dq[1]) % 2 == 0)


# ---- construction of the enlarged-block operators at the same shape ----------------------------------------------
def make_opbuild_case(norbs=40, nelec=40, M=4000, left_sites=None, op_sites=6, other_sites=6, seed=20260, sigma_n=2.2, sigma_s=1.6):
    """Children and product tables of the enlarged left block of make_big_block: (renormalised M-state block) x (one site).  The sector
    tables are those of the full chain position; to bound memory the blocks carry the operator arrays of only `op_sites` own sites
    (+ the dot) and `other_sites` sites of the other block.  Returns (left BlockSpec, dot BlockSpec, operators of the enlarged block,
    product tables, h1, h2); operator values are seeded random numbers (the scatter's throughput does not depend on them)."""
    nl = norbs // 2 if left_sites is None else left_sites
    f = nelec / norbs
    sys = renormalised_sectors(nl - 1, f * (nl - 1), M, sigma_n, sigma_s)
    own = list(range(min(op_sites, nl - 1)))
    dot_site = [nl - 1]
    others = list(range(nl, min(norbs, nl + other_sites)))
    left = make_block(sys, own, dot_site + others, loop=True)
    dot = make_block({(0, 0): 1, (1, 1): 1, (2, 0): 1}, dot_site, own + others, loop=True)
    enlarged = make_block(add_dot(sys), own + dot_site, others, loop=True)
    rng = np.random.default_rng(seed)
    for blk in (left, dot):
        dims = blk.dims.astype(np.int64)
        for op in blk.ops:
            if op.data is None:
                n = int((op.allowed.astype(np.int64) * np.outer(dims, dims)).sum())
                op.data = rng.standard_normal(n) * 0.1
    pt = product_tables(list(sys), np.array(list(sys.values()), np.int32))
    h1 = rng.standard_normal((norbs, norbs)); h1 = h1 + h1.T
    L = rng.standard_normal((norbs * norbs, 4)) * 0.1
    h2 = (L @ L.T).reshape(norbs, norbs, norbs, norbs)
    return left, dot, enlarged.ops, pt, h1, h2


# ---- the same big block with FACTORISED enlarged-block operators (SURVEY.md 7 "hard parts"; option "factorised") ------------------------
def make_child(sectors, sites, ccd_sites, normal, comp_pairs, loop):
    """A child of an enlarged block with the operator arrays the construction of its parent reads: HAM, OVERLAP, CRE_i (own sites),
    CRE_CRE_DESCOMP_i (ccd_sites), CRE_DES / CRE_CRE for its own site pairs (normal) and CRE_DESCOMP / DES_DESCOMP for comp_pairs."""
    q = np.array([[N, s, 0] for (N, s) in sectors], dtype=np.int32)
    dims = np.array(list(sectors.values()), dtype=np.int32)
    blk = BlockSpec(q=q, dims=dims, sites=tuple(sites), loop=loop)
    blk.ops.append(_op(q, HAM, (), 0, (0, 0, 0), False))
    ovl = _op(q, OVERLAP, (), 0, (0, 0, 0), False)
    ovl.data = np.concatenate([np.eye(int(d)).ravel() for d in dims])
    blk.ops.append(ovl)
    for i in sites:
        blk.ops.append(_op(q, CRE, (i,), 0, (1, 1, 0), True))
    for i in ccd_sites:
        blk.ops.append(_op(q, CRE_CRE_DESCOMP, (i,), 0, (1, 1, 0), True))
    if normal:
        for a, i in enumerate(sites):
            for j in sites[:a + 1]:
                for comp, s in ((0, 0), (1, 2)):
                    blk.ops.append(_op(q, CRE_DES, (i, j), comp, (0, s, 0), False))
                    blk.ops.append(_op(q, CRE_CRE, (i, j), comp, (2, s, 0), False))
    for (i, j) in comp_pairs:
        assert i >= j
        for comp, s in ((0, 0), (1, 2)):
            blk.ops.append(_op(q, CRE_DESCOMP, (i, j), comp, (0, s, 0), False))
            blk.ops.append(_op(q, DES_DESCOMP, (i, j), comp, (-2, s, 0), False))
    return blk


def make_product_case(norbs=40, nelec=40, M=4000, left_sites=None, seed=20260, sigma_n=2.2, sigma_s=1.6):
    """Both children of the big block of make_big_block as PRODUCTS (renormalised M-state block) x (one site), described by THEIR children:
    what a sweep holds in memory before the reference's Op::build (or b2d_build_enlarged_op) constructs the enlarged operators.
    Returns dict(left=(child, dot, tables, enlarged), right=(...), h1, h2, dq); operator values are seeded random numbers: the dots get host
    data (their 1 x 1 elements become factors), the M-state children are filled on the device."""
    nl = norbs // 2 if left_sites is None else left_sites
    f = nelec / norbs
    lsites, rsites = list(range(nl)), list(range(nl, norbs))
    s_sites, d_l = lsites[:-1], lsites[-1]          # S' and the system dot
    d_r, e_sites = rsites[0], rsites[1:]            # environment dot and E'
    sys = renormalised_sectors(len(s_sites), f * len(s_sites), M, sigma_n, sigma_s)
    env = renormalised_sectors(len(e_sites), f * len(e_sites), M, sigma_n, sigma_s)
    dot = {(0, 0): 1, (1, 1): 1, (2, 0): 1}
    rng = np.random.default_rng(seed)

    def pairs(a):   # i >= j over a sorted list
        a = sorted(a)
        return [(i, j) for k, i in enumerate(a) for j in a[:k + 1]]
    # system side (normal two-index operators): S' needs the complementary operators that carry the dot index, the dot everything
    s_child = make_child(sys, s_sites, [d_l] + rsites, True, [(I, d_l) for I in rsites] + [(d_l, d_l)], loop=False)
    l_dot = make_child(dot, [d_l], s_sites + rsites, True, [(I, j) for j in s_sites for I in rsites], loop=True)
    # environment side (complementary two-index operators for the pairs of the left block)
    e_child = make_child(env, e_sites, lsites + [d_r], False, pairs(lsites + [d_r]), loop=False)
    r_dot = make_child(dot, [d_r], lsites + e_sites, True, pairs(lsites) + [(j, I) for j in e_sites for I in lsites], loop=True)
    for blk in (l_dot, r_dot):
        dims = blk.dims.astype(np.int64)
        for op in blk.ops:
            if op.data is None:
                n = int((op.allowed.astype(np.int64) * np.outer(dims, dims)).sum())
                op.data = rng.standard_normal(n) * 0.3
    lt = product_tables(list(sys), np.array(list(sys.values()), np.int32))
    rt = product_tables(list(env), np.array(list(env.values()), np.int32))
    left = make_block({(int(k[0]), int(k[1])): int(d) for k, d in zip(lt["q"], lt["dims"])}, lsites, rsites, loop=True)
    right = make_block({(int(k[0]), int(k[1])): int(d) for k, d in zip(rt["q"], rt["dims"])}, rsites, lsites, loop=False)
    h1 = rng.standard_normal((norbs, norbs)); h1 = h1 + h1.T
    Lc = rng.standard_normal((norbs * norbs, 4)) * 0.1
    h2 = (Lc @ Lc.T).reshape(norbs, norbs, norbs, norbs)
    return dict(left=(s_child, l_dot, lt, left), right=(e_child, r_dot, rt, right), h1=h1, h2=h2, dq=(nelec, 0, 0), norbs=norbs, seed=seed)


def make_big_block_from_products(case, device=0, options=None, factorised=True, rank=0, nranks=1):
    """SpinBlock whose two children are built ON THE DEVICE from the case's grandchildren: factorised (no enlarged operator is ever
    materialised) or, for comparison at small M, materialised by the scatter kernel."""
    opts = dict(options or {})
    opts["factorised"] = 1 if factorised else 0
    return SpinBlock.from_products(case["left"], case["right"], case["dq"], norbs=case["norbs"], device=device, options=opts, rank=rank, nranks=nranks,
                                   integrals=(case["h1"], case["h2"], np.zeros(case["norbs"], np.int32)), fill_seed=case["seed"])
