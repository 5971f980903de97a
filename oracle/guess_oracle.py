"""CPU restatement of the guess-wavefunction transform of a two-dot step.  TEST INFRASTRUCTURE ONLY: imported by
tests/ only, never by the product (block_b200/).  PARITY PINNED: tests/test_guess_oracle.py checks it against the trial vectors of the
real reference (tests/golden/guess_*.npz, dumped by oracle/ref_dump.cpp from oracle/_ref/block_dump).

Follows GuessWave::transform_previous_wavefunction, two-dot branch (guess_wavefunction.C:524-636):

    previous wavefunction  [S (x) d1] [E_old (x) d2]   (rows: old left sectors, columns: collected E_old (x) d2 sectors)
      1. TransformLeftBlock   (:17-31)    rows  -> S' = the renormalised system block:  L_q^T psi[q, b]
      2. onedot_shufflesysdot (:434-485, :200-256)   [S'][E_old d2] -> [S' d2][E_old]: columns un-collected, fermion/recoupling sign
         (getCommuteParity), spin recoupling 6j, rows collected
      3. TransformRightBlock  (:33-50)    columns E_old -> the un-truncated basis of the new environment side:  psi[a, c] R_c^T

A wavefunction is a dict {(row sector, column sector): ndarray}; a StateInfo is the dict of tables the dump holds ("q", "dims",
"new_quanta_map", "unc.q", "unc.dims", "unc.lmap", "unc.rmap", "old_to_new", "old_to_new_begin").
"""
import math

import numpy as np

from oracle import dmrg_oracle as O


def stateinfo(rec, prefix):
    """The tables of one dumped StateInfo (oracle/ref_dump.cpp dump_si_tables)."""
    out = {}
    for k in rec.keys() if hasattr(rec, "keys") else rec.files:
        if k.startswith(prefix):
            out[k[len(prefix):]] = np.asarray(rec[k])
    return out


def allow(dq, ql, qr):
    """SpinQuantum::allow (SpinQuantum.C:99-107): dq in ql (+) qr, abelian point group."""
    return O.qn_allow(tuple(int(x) for x in dq), tuple(int(x) for x in ql), tuple(int(x) for x in qr))


def unpack_blocks(allowed, data, row_dims, col_dims):
    """Row-major list of allowed blocks -> {(i, j): matrix}."""
    w, off = {}, 0
    for i in range(allowed.shape[0]):
        for j in range(allowed.shape[1]):
            if allowed[i, j]:
                n = int(row_dims[i]) * int(col_dims[j])
                w[(i, j)] = np.array(data[off:off + n]).reshape(int(row_dims[i]), int(col_dims[j]))
                off += n
    assert off == data.size
    return w


def unpack_rotation(shape, data):
    """vector<Matrix> of a Rotation file: entry q is d_q x m_q, empty for a dropped sector."""
    out, off = [], 0
    for r, c in shape:
        r, c = int(r), int(c)
        out.append(np.array(data[off:off + r * c]).reshape(r, c) if c else None)
        off += r * c if c else 0
    return out


def flatten(w, nrows, ncols, dq, row_q, col_q, row_dims, col_dims):
    """Wavefunction::FlattenInto order (wavefunction.C:167-186): allowed blocks row-major, each row-major."""
    parts = []
    for i in range(nrows):
        for j in range(ncols):
            if allow(dq, row_q[i], col_q[j]):
                parts.append(w.get((i, j), np.zeros((int(row_dims[i]), int(col_dims[j])))).ravel())
    return np.concatenate(parts) if parts else np.zeros(0)


def commute_parity(a, b, c):
    """getCommuteParity (BaseOperator.C:20-53), spin-adapted, abelian point group."""
    return O.commute_parity(tuple(int(x) for x in a), tuple(int(x) for x in b), tuple(int(x) for x in c))


def transform_left_block(old, sys, oldright, lrot, dq):
    """TransformLeftBlock (guess_wavefunction.C:17-31): tempoldWave(a, b) = L[olda]^T . oldWave(olda, b)."""
    out = {}
    for a in range(len(sys["dims"])):
        olda = int(sys["new_quanta_map"][a])
        for b in range(len(oldright["dims"])):
            if (olda, b) in old:
                assert allow(dq, sys["q"][a], oldright["q"][b])
                out[(a, b)] = lrot[olda].T @ old[(olda, b)]
    return out


def shuffle_sysdot(temp, sys, dot, env, oldright, left, dq):
    """onedot_shufflesysdot (guess_wavefunction.C:434-441) = onedot_twoindex_to_threeindex_shufflesysdot (:443-485) followed by
    onedot_threeindex_to_twoindex_wavefunction (:200-256): [S'][E_old d] -> [S' d][E_old]."""
    nsys, ndot, nenv = len(sys["dims"]), len(dot["dims"]), len(env["dims"])
    # UnCollectQuantaAlongColumns (wavefunction.C:348-375): column pieces of each collected sector in oldToNewState order
    three = {}   # (a, b, c) -> {uncollected column index of oldright: block}
    for bcol in range(len(oldright["dims"])):
        first = 0
        for k in range(int(oldright["old_to_new_begin"][bcol]), int(oldright["old_to_new_begin"][bcol + 1])):
            u = int(oldright["old_to_new"][k])
            size = int(oldright["unc.dims"][u])
            for a in range(nsys):
                if (a, bcol) in temp and allow(dq, sys["q"][a], oldright["unc.q"][u]):
                    b = int(oldright["unc.rmap"][u])   # the dot sector
                    c = int(oldright["unc.lmap"][u])   # the E_old sector
                    parity = commute_parity(env["q"][c], dot["q"][b], oldright["unc.q"][u])
                    three.setdefault((a, b, c), {})[u] = parity * temp[(a, bcol)][:, first:first + size]
            first += size
    # three-index -> two-index (:200-253): rows (a, b) of the un-collected S' (x) dot, spin recoupling
    two = {}
    J = int(dq[1])
    for (a, b, c), slots in three.items():
        for ab in range(len(left["unc.dims"])):
            if int(left["unc.lmap"][ab]) != a or int(left["unc.rmap"][ab]) != b:
                continue
            if not allow(dq, left["unc.q"][ab], env["q"][c]):
                continue
            A, B, AB, C = int(sys["q"][a][1]), int(dot["q"][b][1]), int(left["unc.q"][ab][1]), int(env["q"][c][1])
            for cb, block in slots.items():    # prevUnCollectedSI.quantaMap(c, b): the un-collected E_old (x) dot sectors
                CB = int(oldright["unc.q"][cb][1])
                scale = O.six_j(A, B, AB, C, J, CB) * math.sqrt((AB + 1.0) * (CB + 1.0)) * (-1.0) ** int((A + B + J + C) / 2)
                # Symmetry::spatial_sixj (Symmetry.C:520-526), abelian: 1 if the irreps couple as the sectors say, else 0
                Al, Bl, ABl, Cl, Jl, CBl = int(sys["q"][a][2]), int(dot["q"][b][2]), int(left["unc.q"][ab][2]), int(env["q"][c][2]), int(dq[2]), int(oldright["unc.q"][cb][2])
                if ABl != (Al ^ Bl) or CBl != (Bl ^ Cl) or Jl != (ABl ^ Cl):
                    scale = 0.0
                tgt = two.setdefault((ab, c), np.zeros((int(left["unc.dims"][ab]), int(env["dims"][c]))))
                d_dot = int(dot["dims"][b])
                assert d_dot == 1, "dot sectors of a spin-adapted site hold one state"
                tgt += scale * block
    # CollectQuantaAlongRows (wavefunction.C:245-270): un-collected row pieces stacked in oldToNewState order
    out = {}
    for lq in range(len(left["dims"])):
        for c in range(nenv):
            if not allow(dq, left["q"][lq], env["q"][c]):
                continue
            m = np.zeros((int(left["dims"][lq]), int(env["dims"][c])))
            first = 0
            for k in range(int(left["old_to_new_begin"][lq]), int(left["old_to_new_begin"][lq + 1])):
                ab = int(left["old_to_new"][k])
                size = int(left["unc.dims"][ab])
                if (ab, c) in two:
                    m[first:first + size, :] = two[(ab, c)]
                first += size
            assert first == m.shape[0]
            out[(lq, c)] = m
    return out


def transform_right_block(tempnew, env, right, rrot, dq, left):
    """TransformRightBlock (guess_wavefunction.C:33-50): trial(a, transB) = tempnewWave(a, b) . R[transB]^T."""
    out = {}
    for (a, b), m in tempnew.items():
        tb = int(env["new_quanta_map"][b])
        assert allow(dq, left["q"][a], right["q"][tb])
        out[(a, tb)] = out.get((a, tb), 0) + m @ rrot[tb].T
    return out


def transform_previous_wavefunction(rec, root):
    """The trial vector of root `root`, flat in FlattenInto order, from one dumped record."""
    p = "gw%d." % root
    dq = rec[p + "dq"][:3]
    sys, dot, left, right = (stateinfo(rec, p + n + ".") for n in ("sys", "dot", "left", "right"))
    oldleft, oldright, env = (stateinfo(rec, p + n + ".") for n in ("oldleft", "oldright", "env"))
    old = unpack_blocks(rec[p + "old.allowed"], rec[p + "old.data"], oldleft["dims"], oldright["dims"])
    lrot = unpack_rotation(rec[p + "lrot.shape"], rec[p + "lrot.data"])
    rrot = unpack_rotation(rec[p + "rrot.shape"], rec[p + "rrot.data"])
    t1 = transform_left_block(old, sys, oldright, lrot, dq)
    t2 = shuffle_sysdot(t1, sys, dot, env, oldright, left, dq)
    t3 = transform_right_block(t2, env, right, rrot, dq, left)
    return flatten(t3, len(left["dims"]), len(right["dims"]), dq, left["q"], right["q"], left["dims"], right["dims"])


# ---- one-dot branch: GuessWave::onedot_transform_wavefunction (guess_wavefunction.C:832-936) ---------------------------------------
def onedot_rotate(old, oldleft, oldcol, rows, cols_dims, lrot, rrot):
    """:870-911.  tmp(a, transC) = old(a, c) . R[transC]^T for every block of the previous wavefunction (transC = newQuantaMap of
    its column sector), then new(a', c) = L[oldA]^T . tmp(oldA, c) with oldA = newQuantaMap of the new row sector a'."""
    ncol = len(cols_dims)
    tmp = {}
    for (a, c), m in old.items():
        tc = int(oldcol["new_quanta_map"][c])
        tmp[(a, tc)] = m @ rrot[tc].T
    out = {}
    for c in range(ncol):
        for a in range(len(rows["dims"])):
            olda = int(rows["new_quanta_map"][a])
            if (olda, c) in tmp:
                out[(a, c)] = lrot[olda].T @ tmp[(olda, c)]
    return out


def transform_previous_wavefunction_onedot(rec, root):
    """The one-dot trial vector of root `root`, flat in FlattenInto order.  transpose_guess_wave (dot on the system side): the rotated
    wavefunction is in [S'][E'.dot] form and the dot is moved to the system by the same shuffle as in the two-dot case."""
    p = "gw%d." % root
    dq = rec[p + "dq"][:3]
    transpose = int(rec["gw.nroots"][1]) != 0
    left, right, oldleft, oldcol = (stateinfo(rec, p + n + ".") for n in ("left", "right", "oldleft", "oldcol"))
    old = unpack_blocks(rec[p + "old.allowed"], rec[p + "old.data"], oldleft["dims"], oldcol["dims"])
    lrot = unpack_rotation(rec[p + "lrot.shape"], rec[p + "lrot.data"])
    rrot = unpack_rotation(rec[p + "rrot.shape"], rec[p + "rrot.data"])
    if not transpose:
        t = onedot_rotate(old, oldleft, oldcol, left, right["dims"], lrot, rrot)
        return flatten(t, len(left["dims"]), len(right["dims"]), dq, left["q"], right["q"], left["dims"], right["dims"])
    sys, dot, newenv = (stateinfo(rec, p + n + ".") for n in ("sys", "dot", "newenv"))
    t1 = onedot_rotate(old, oldleft, oldcol, sys, newenv["dims"], lrot, rrot)          # [S'][E'.dot]
    t2 = shuffle_sysdot(t1, sys, dot, right, newenv, left, dq)                          # -> [S'.dot][E']
    return flatten(t2, len(left["dims"]), len(right["dims"]), dq, left["q"], right["q"], left["dims"], right["dims"])


# ---- first block iteration of a sweep: GuessWave::transpose_previous_wavefunction (guess_wavefunction.C:55-84), two-dot ------------
def transpose_previous_wavefunction(rec, root):
    """trial(i, j) = getCommuteParity(q_right_old[i], q_left_old[j], dq) . old(j, i)^T: the previous wavefunction seen from the other end
    of the chain (a transposition, not a Hermitian conjugate, with the fermionic / recoupling sign of swapping the two blocks)."""
    p = "gw%d." % root
    dq = rec[p + "dq"][:3]
    left, right, oldleft, oldcol = (stateinfo(rec, p + n + ".") for n in ("left", "right", "oldleft", "oldcol"))
    old = unpack_blocks(rec[p + "old.allowed"], rec[p + "old.data"], oldleft["dims"], oldcol["dims"])
    out = {}
    for i in range(len(left["dims"])):
        for j in range(len(right["dims"])):
            if allow(dq, left["q"][i], right["q"][j]):
                assert (j, i) in old
                out[(i, j)] = commute_parity(oldcol["q"][i], oldleft["q"][j], dq) * old[(j, i)].T
    return flatten(out, len(left["dims"]), len(right["dims"]), dq, left["q"], right["q"], left["dims"], right["dims"])


# ---- first block iteration of a ONE-DOT sweep: onedot_transpose_wavefunction (guess_wavefunction.C:140-198, :200-256, :402-432) -----
def onedot_transpose_wavefunction(rec, root):
    """[S.d][E] -> [E.d][S]: rows of the previous wavefunction un-collected into (S sector, dot sector) pieces (:402-432), every piece
    transposed with the sign of moving |S>|d> past |E> (:163-190), then re-coupled into the rows (E, d) of the new left block with the
    same 6j coefficient as the shuffle (:200-256)."""
    p = "gw%d." % root
    dq = rec[p + "dq"][:3]
    left, sys, dot, right = (stateinfo(rec, p + n + ".") for n in ("left", "sys", "dot", "right"))
    oldleft, oldsys, oldcol = (stateinfo(rec, p + n + ".") for n in ("oldleft", "oldsys", "oldcol"))
    old = unpack_blocks(rec[p + "old.allowed"], rec[p + "old.data"], oldleft["dims"], oldcol["dims"])
    J = int(dq[1])
    # un-collect the rows of the previous wavefunction: piece u = (a: S sector, b: dot sector, quantum AB)
    pieces = {}
    for lq in range(len(oldleft["dims"])):
        first = 0
        for k in range(int(oldleft["old_to_new_begin"][lq]), int(oldleft["old_to_new_begin"][lq + 1])):
            u = int(oldleft["old_to_new"][k])
            size = int(oldleft["unc.dims"][u])
            for c in range(len(oldcol["dims"])):
                if (lq, c) in old and allow(dq, oldleft["unc.q"][u], oldcol["q"][c]):
                    pieces[(u, c)] = old[(lq, c)][first:first + size, :]
            first += size
    two = {}
    for (u, c), blk in pieces.items():
        a, b = int(oldleft["unc.lmap"][u]), int(oldleft["unc.rmap"][u])
        Aq, Bq, ABq, Cq = oldsys["q"][a], dot["q"][b], oldleft["unc.q"][u], oldcol["q"][c]
        parity = commute_parity(Aq, Bq, ABq) * commute_parity(ABq, Cq, dq)
        tr = parity * blk.T                                  # (E sector c) x (S sector a)
        # new roles: rows (E = sys sector c, dot b), columns S = right sector a
        for v in range(len(left["unc.dims"])):
            if int(left["unc.lmap"][v]) != c or int(left["unc.rmap"][v]) != b:
                continue
            if not allow(dq, left["unc.q"][v], right["q"][a]):
                continue
            A, B, AB, C, CB = int(sys["q"][c][1]), int(dot["q"][b][1]), int(left["unc.q"][v][1]), int(right["q"][a][1]), int(ABq[1])
            scale = O.six_j(A, B, AB, C, J, CB) * math.sqrt((AB + 1.0) * (CB + 1.0)) * (-1.0) ** int((A + B + J + C) / 2)
            Al, Bl, ABl, Cl, CBl = int(sys["q"][c][2]), int(dot["q"][b][2]), int(left["unc.q"][v][2]), int(right["q"][a][2]), int(ABq[2])
            if ABl != (Al ^ Bl) or CBl != (Bl ^ Cl) or int(dq[2]) != (ABl ^ Cl):
                scale = 0.0
            tgt = two.setdefault((v, a), np.zeros((int(left["unc.dims"][v]), int(right["dims"][a]))))
            tgt += scale * tr
    out = {}
    for lq in range(len(left["dims"])):
        for a in range(len(right["dims"])):
            if not allow(dq, left["q"][lq], right["q"][a]):
                continue
            m = np.zeros((int(left["dims"][lq]), int(right["dims"][a])))
            first = 0
            for k in range(int(left["old_to_new_begin"][lq]), int(left["old_to_new_begin"][lq + 1])):
                v = int(left["old_to_new"][k])
                size = int(left["unc.dims"][v])
                if (v, a) in two:
                    m[first:first + size, :] = two[(v, a)]
                first += size
            out[(lq, a)] = m
    return flatten(out, len(left["dims"]), len(right["dims"]), dq, left["q"], right["q"], left["dims"], right["dims"])
