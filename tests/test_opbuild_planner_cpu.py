"""The C++ host planner of enlarged-block operator construction (block_b200/csrc/opbuild.hpp, through b2d_enlarged_op_products on a
planning-only context: integer / scalar work, no device) against the oracle's restatement of the reference's Op::build
(oracle/opbuild_oracle.py, itself pinned against the real reference): the SAME products of child operators with the SAME scalar
factors for every operator of every dumped block iteration - Transposeview flags, commute parities, 6j recoupling factors and
integral-weighted complementary factors included."""
import glob
import os

import numpy as np
import pytest

from block_b200 import hotpath
from oracle import dumpio
from oracle import opbuild_oracle as B

FIXTURES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "opbuild_*.npz")))


def oracle_products(monkeypatch, pi, op, ints, hubbard, lid, rid):
    calls = []

    def rec_product(pi_, a, b, a_on_left, c, scale=1.0):
        l, r = (a, b) if a_on_left else (b, a)
        calls.append((lid[id(l.op)], bool(l.t), rid[id(r.op)], bool(r.t), float(scale)))

    def rec_trace(pi_, a, a_on_left, c, scale=1.0):
        calls.append((lid[id(a.op)], bool(a.t), None, False, float(scale)) if a_on_left else (None, False, rid[id(a.op)], bool(a.t), float(scale)))

    monkeypatch.setattr(B, "tensor_product", rec_product)
    monkeypatch.setattr(B, "tensor_trace", rec_trace)
    B.build_operator(pi, op, ints, hubbard)
    monkeypatch.undo()
    return calls


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(f)[:-4] for f in FIXTURES])
def test_planner_lists_the_reference_products(monkeypatch, path):
    rec = dict(np.load(path))
    pi, ref, ints = B.ProductInfo.from_record(rec), dumpio.block_from(rec, "LA."), B.Integrals.from_record(rec)
    hubbard = int(rec["meta"][7]) == B.O.HUBBARD_HAM
    left, right = hotpath.block_spec_from_record(rec, "LL."), hotpath.block_spec_from_record(rec, "LR.")
    pb = hotpath.ProductBlock(left, right, pi.q, pi.dims, pi.lmap, pi.rmap, pi.unc_dims, pi.old_to_new, device=-1)   # planning only
    pb.set_integrals(ints.h1, ints.h2, ints.irreps, ints.one_tol, ints.two_tol)
    lid = {id(op): k for k, op in enumerate(pi.left.ops)}
    rid = {id(op): k for k, op in enumerate(pi.right.ops)}
    key = lambda c: (-1 if c[0] is None else c[0], c[1], -1 if c[2] is None else c[2], c[3])
    try:
        total, by_type = 0, {}
        for op in ref.ops:
            want = sorted(oracle_products(monkeypatch, pi, op, ints, hubbard, lid, rid), key=key)
            got = sorted(pb.products(op.optype, op.orbs, op.dq, hubbard), key=key)
            assert [key(c) for c in got] == [key(c) for c in want], (op.optype, op.orbs, op.comp)
            for g, w in zip(got, want):
                assert abs(g[4] - w[4]) <= 1e-12 * max(1.0, abs(w[4])), (op.optype, op.orbs, op.comp, g, w)
            total += len(want)
            by_type[op.optype] = by_type.get(op.optype, 0) + len(want)
        assert total >= len(ref.ops)
        assert by_type.get(B.HAM, 0) >= 2 and by_type.get(B.CRE_CRE_DESCOMP, 0) >= 1, by_type
    finally:
        pb.close()


def test_planner_needs_integrals_for_complementary_operators():
    rec = dict(np.load(FIXTURES[0]))
    pi = B.ProductInfo.from_record(rec)
    left, right = hotpath.block_spec_from_record(rec, "LL."), hotpath.block_spec_from_record(rec, "LR.")
    pb = hotpath.ProductBlock(left, right, pi.q, pi.dims, pi.lmap, pi.rmap, pi.unc_dims, pi.old_to_new, device=-1)
    try:
        ref = dumpio.block_from(rec, "LA.")
        comp = next(op for op in ref.ops if op.optype == B.CRE_DESCOMP)
        with pytest.raises(hotpath.B2DError, match="integrals"):
            pb.products(comp.optype, comp.orbs, comp.dq)
        with pytest.raises(hotpath.B2DError, match="no CUDA device"):
            pb.build(B.OVERLAP, (), (0, 0, 0), False)       # building needs the device: no CPU fallback
    finally:
        pb.close()
