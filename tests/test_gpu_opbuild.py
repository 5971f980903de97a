"""First device step of SURVEY.md N2: the TensorProduct / TensorTrace scatter (operatorfunctions.C:19-254) on the GPU, through the C ABI
(b2d_set_product_stateinfo / b2d_product_op_create / b2d_product_op_accumulate).  WHICH child products enter an operator and their
integral factors are taken from the oracle's restatement of the reference's Op::build (oracle/opbuild_oracle.py, pinned against the
real reference); the device performs every product; the result must equal the enlarged-block operators the REAL reference built
(tests/golden/opbuild_*.npz) - every operator type of an energy sweep, Hamiltonian and complementary operators included."""
import glob
import os

import numpy as np
import pytest

from block_b200 import hotpath
from oracle import dumpio
from oracle import opbuild_oracle as B

pytestmark = pytest.mark.gpu
FIXTURES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "opbuild_*.npz")))


def recorded_products(monkeypatch, pi, op, ints, hubbard):
    """Run the oracle's build of `op` with its two primitives replaced by recorders: the list of (left op, left transposed, right op,
    right transposed, scale) products; None = identity on that child (TensorTrace)."""
    calls = []

    def rec_product(pi_, a, b, a_on_left, c, scale=1.0):
        l, r = (a, b) if a_on_left else (b, a)
        calls.append((l.op, l.t, r.op, r.t, scale))

    def rec_trace(pi_, a, a_on_left, c, scale=1.0):
        calls.append((a.op, a.t, None, False, scale) if a_on_left else (None, False, a.op, a.t, scale))

    monkeypatch.setattr(B, "tensor_product", rec_product)
    monkeypatch.setattr(B, "tensor_trace", rec_trace)
    B.build_operator(pi, op, ints, hubbard)
    monkeypatch.undo()
    return calls


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(f)[:-4] for f in FIXTURES])
def test_device_tensor_products_rebuild_the_reference_operators(monkeypatch, path):
    rec = dict(np.load(path))
    pi, ref, ints = B.ProductInfo.from_record(rec), dumpio.block_from(rec, "LA."), B.Integrals.from_record(rec)
    hubbard = int(rec["meta"][7]) == B.O.HUBBARD_HAM
    left, right = hotpath.block_spec_from_record(rec, "LL."), hotpath.block_spec_from_record(rec, "LR.")
    pb = hotpath.ProductBlock(left, right, pi.q, pi.dims, pi.lmap, pi.rmap, pi.unc_dims, pi.old_to_new, device=0)
    lid = {id(op): k for k, op in enumerate(pi.left.ops)}
    rid = {id(op): k for k, op in enumerate(pi.right.ops)}
    try:
        built, products = {}, 0
        for op in ref.ops:
            calls = recorded_products(monkeypatch, pi, op, ints, hubbard)
            pid = pb.create(op.dq, op.fermion)
            for lop, lt, rop, rt, scale in calls:
                pb.accumulate(pid, None if lop is None else lid[id(lop)], None if rop is None else rid[id(rop)], lt, rt, scale)
            products += len(calls)
            allowed, data = pb.download(pid)
            assert np.array_equal(allowed, op.allowed), (op.optype, op.orbs, op.comp)
            off = 0
            for i in range(len(pi.dims)):
                for j in range(len(pi.dims)):
                    if not op.allowed[i, j]:
                        continue
                    blk = op.blocks[(i, j)]
                    got = data[off:off + blk.size].reshape(blk.shape)
                    off += blk.size
                    scale = max(1.0, float(np.abs(blk).max()))
                    assert np.abs(got - blk).max() <= 1e-12 * scale, (op.optype, op.orbs, op.comp, (i, j), np.abs(got - blk).max())
            assert off == data.size
            built[op.optype] = built.get(op.optype, 0) + 1
        assert B.HAM in built and B.CRE in built and B.CRE_CRE_DESCOMP in built, built
        assert products >= len(ref.ops)
        assert pb.kernel_launches() >= products          # every product ran as a device launch
    finally:
        pb.close()


@pytest.mark.parametrize("factorised", [0, 1], ids=["materialised", "factorised"])
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(f)[:-4] for f in FIXTURES])
def test_build_enlarged_operators_without_the_oracle(path, factorised):
    """b2d_build_enlarged_op: the library's own host planner (block_b200/csrc/opbuild.hpp) decides the products and the integral
    factors, the device performs them; inputs are only the two children, the product StateInfo and the integrals.  Every
    operator of the enlarged block must equal the one the REAL reference built."""
    rec = dict(np.load(path))
    pi, ref, ints = B.ProductInfo.from_record(rec), dumpio.block_from(rec, "LA."), B.Integrals.from_record(rec)
    hubbard = int(rec["meta"][7]) == B.O.HUBBARD_HAM
    left, right = hotpath.block_spec_from_record(rec, "LL."), hotpath.block_spec_from_record(rec, "LR.")
    pb = hotpath.ProductBlock(left, right, pi.q, pi.dims, pi.lmap, pi.rmap, pi.unc_dims, pi.old_to_new, device=0, options={"factorised": factorised})
    pb.set_integrals(ints.h1, ints.h2, ints.irreps, ints.one_tol, ints.two_tol)
    try:
        worst = 0.0
        for op in ref.ops:
            pid = pb.build(op.optype, op.orbs, op.dq, op.fermion, hubbard)
            allowed, data = pb.download(pid)
            assert np.array_equal(allowed, op.allowed), (op.optype, op.orbs, op.comp)
            off = 0
            for i in range(len(pi.dims)):
                for j in range(len(pi.dims)):
                    if not op.allowed[i, j]:
                        continue
                    blk = op.blocks[(i, j)]
                    got = data[off:off + blk.size].reshape(blk.shape)
                    off += blk.size
                    err = np.abs(got - blk).max() / max(1.0, float(np.abs(blk).max()))
                    worst = max(worst, err)
                    assert err <= 1e-12, (op.optype, op.orbs, op.comp, (i, j), err)
        assert pb.kernel_launches() > len(ref.ops)
        print("%s: %d operators rebuilt on the device, worst relative difference %.1e" % (os.path.basename(path), len(ref.ops), worst))
    finally:
        pb.close()
