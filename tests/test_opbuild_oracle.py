"""The operator-construction oracle (oracle/opbuild_oracle.py) against the real reference: for block iterations
dumped by oracle/_ref/block_dump (tests/golden/opbuild_*.npz) every CRE, CRE_CRE, CRE_DES and OVERLAP operator of the enlarged
block, rebuilt from the two children with the restated TensorProduct / TensorTrace, must equal what the reference's own Op::build
produced - block for block, allowed mask included."""
import glob
import os

import numpy as np
import pytest

from oracle import dumpio
from oracle import opbuild_oracle as B

FIXTURES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "opbuild_*.npz")))


@pytest.fixture(scope="module", params=FIXTURES, ids=[os.path.basename(f)[:-4] for f in FIXTURES])
def record(request):
    rec = dict(np.load(request.param))
    return rec, B.ProductInfo.from_record(rec), dumpio.block_from(rec, "LA.")


def test_fixtures_exist():
    assert len(FIXTURES) >= 3


def test_product_stateinfo_is_consistent(record):
    rec, pi, ref = record
    # every collected sector is the concatenation of its uncollected (left sector, right sector) pieces
    for cq, pieces in enumerate(pi.old_to_new):
        assert sum(int(pi.unc_dims[p]) for p in pieces) == int(pi.dims[cq])
        for p in pieces:
            assert int(pi.unc_dims[p]) == int(pi.left.dims[pi.lmap[p]]) * int(pi.right.dims[pi.rmap[p]])
            lq, rq = pi.left.q[pi.lmap[p]], pi.right.q[pi.rmap[p]]
            assert B.O.qn_allow(tuple(pi.q[cq]), tuple(lq), tuple(rq))


def test_normal_operators_match_reference(record):
    rec, pi, ref = record
    checked = {B.CRE: 0, B.CRE_CRE: 0, B.CRE_DES: 0, B.OVERLAP: 0}
    for op in ref.ops:
        if op.optype not in checked:
            continue
        mine = B.build_normal_operator(pi, op)
        assert np.array_equal(mine.allowed, op.allowed), (op.optype, op.orbs, op.comp)
        for key, blk in op.blocks.items():
            scale = max(1.0, float(np.abs(blk).max()))
            assert np.abs(mine.blocks[key] - blk).max() <= 1e-13 * scale, (op.optype, op.orbs, op.comp, key)
        checked[op.optype] += 1
    assert checked[B.CRE] >= 1
    assert sum(checked.values()) >= 3, checked


def test_fixtures_exercise_cross_child_products():
    """At least one fixture holds two-index operators with one index on each child (a x b products, not only O x 1)."""
    found = 0
    for f in FIXTURES:
        rec = dict(np.load(f))
        pi, ref = B.ProductInfo.from_record(rec), dumpio.block_from(rec, "LA.")
        lsites = set(pi.left.sites)
        found += sum(1 for op in ref.ops if op.optype in (B.CRE_CRE, B.CRE_DES) and len({o in lsites for o in op.orbs}) == 2)
    assert found >= 4


def test_two_index_complementary_operators_match_reference(record):
    """CreDesComp / DesDesComp: integral-weighted sums of child products (TensorOp coupling + calcCompfactor restated)."""
    rec, pi, ref = record
    ints = B.Integrals.from_record(rec)
    n = 0
    for op in ref.ops:
        if op.optype not in (B.CRE_DESCOMP, B.DES_DESCOMP):
            continue
        mine = B.build_operator(pi, op, ints)
        assert np.array_equal(mine.allowed, op.allowed), (op.optype, op.orbs, op.comp)
        for key, blk in op.blocks.items():
            scale = max(1.0, float(np.abs(blk).max()))
            assert np.abs(mine.blocks[key] - blk).max() <= 1e-12 * scale, (op.optype, op.orbs, op.comp, key, np.abs(mine.blocks[key] - blk).max())
        n += 1
    if rec["LA.nops"][0] > 10:      # the Hubbard fixture carries no two-index operators
        assert n >= 4


def test_three_index_complementary_and_hamiltonian_match_reference(record):
    """CreCreDesComp (recoupled products with the other child's two-index complementary operators) and Ham of the enlarged block."""
    rec, pi, ref = record
    ints = B.Integrals.from_record(rec)
    hubbard = int(rec["meta"][7]) == B.O.HUBBARD_HAM
    n = {B.CRE_CRE_DESCOMP: 0, B.HAM: 0}
    for op in ref.ops:
        if op.optype not in n:
            continue
        mine = B.build_operator(pi, op, ints, hubbard)
        assert np.array_equal(mine.allowed, op.allowed), (op.optype, op.orbs)
        for key, blk in op.blocks.items():
            scale = max(1.0, float(np.abs(blk).max()))
            err = np.abs(mine.blocks[key] - blk).max()
            assert err <= 1e-12 * scale, (op.optype, op.orbs, key, err)
        n[op.optype] += 1
    assert n[B.HAM] == 1 and n[B.CRE_CRE_DESCOMP] >= 1, n


def test_every_operator_of_the_enlarged_block_is_restated(record):
    rec, pi, ref = record
    types = {op.optype for op in ref.ops}
    assert types <= {B.HAM, B.CRE, B.CRE_CRE, B.DES_DESCOMP, B.CRE_DES, B.CRE_DESCOMP, B.CRE_CRE_DESCOMP, B.OVERLAP}, types
