"""The synthetic benchmark block (block_b200/synthetic.py, bench.py) against the REAL reference (oracle/_ref/ref_bench =
the unmodified reference objects + our driver) and the numpy oracle, at sizes the CPU finishes in seconds."""
import argparse

import numpy as np
import pytest

import bench
from block_b200 import synthetic as S
from oracle import dmrg_oracle as O
from oracle import refbench

needs_ref = pytest.mark.skipif(not refbench.available(), reason="oracle/_ref/ref_bench not built (make -C oracle ref)")


def args(M, left_sites=6, norbs=12, nelec=12):
    return argparse.Namespace(norbs=norbs, nelec=nelec, M=M, left_sites=left_sites)


def counter_values(seed, amp, n):
    """numpy copy of the CUDA library's counter-based stream (kernels.cu counter_uniform)."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * (np.arange(n, dtype=np.uint64) + np.uint64(1))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    return 2.0 * amp * ((z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) - 0.5)


def materialise(blk, op, seed, amp):
    n = int(sum(int(blk.dims[i]) * int(blk.dims[j]) for i, j in zip(*np.nonzero(op.allowed))))
    vals, off = counter_values(seed, amp, n), 0
    op.blocks.clear()
    for i, j in zip(*np.nonzero(op.allowed)):
        k = int(blk.dims[i]) * int(blk.dims[j])
        op.blocks[(int(i), int(j))] = vals[off:off + k].reshape(int(blk.dims[i]), int(blk.dims[j]))
        off += k


def test_sector_model_is_consistent():
    sec = S.renormalised_sectors(9, 9.0, 500)
    assert sum(sec.values()) == 500 and all(d > 0 for d in sec.values())
    big = S.add_dot(sec)
    assert sum(big.values()) >= 3 * 500 and list(big) == sorted(big)
    assert S.csf_count(2, 2, 0) == 3 and S.csf_count(2, 2, 2) == 1 and S.csf_count(3, 3, 1) == 8


@needs_ref
def test_term_flops_planner_equals_reference():
    """The dgemm flops the reference's TensorMultiply issues over ALL multiplyH terms (its own allocate() rule and loops)
    equal the planner's algorithmic flop count: same sectors, same allowed blocks, same term list."""
    a = args(60)
    big, terms = bench._oracle_big(a)
    _, flops, _ = refbench.run(big, terms, list(range(len(terms))), cores=2)
    sb = S.make_big_block(norbs=a.norbs, nelec=a.nelec, M=a.M, left_sites=a.left_sites, device=-1)
    try:
        assert len(terms) == len(sb.terms(all_ranks=True)[0])
        assert flops == sb.sigma_flops(all_ranks=True)
    finally:
        sb.close()


@needs_ref
def test_oracle_tensor_multiply_equals_reference_on_synthetic_block():
    """numpy oracle vs the real reference on the synthetic sectors with the shared counter-based operator values."""
    a = args(80)
    big, terms = bench._oracle_big(a)
    pair = [i for i in range(len(terms)) if terms[i][0].op.optype not in (S.HAM, S.OVERLAP) and terms[i][1].op.optype not in (S.HAM, S.OVERLAP)]
    pick = [pair[int(j)] for j in np.unique(np.linspace(0, len(pair) - 1, 24).astype(int))]
    psi = np.random.default_rng(3).standard_normal(big.size)
    fills = {}
    v = big.zeros()
    c = big.unflatten(psi)
    for i in pick:
        lop, rop, scale = terms[i]
        kl = next(k for k, o in enumerate(big.left.ops) if o is lop.op)
        kr = next(k for k, o in enumerate(big.right.ops) if o is rop.op)
        ls, la = S.fill_params(big.left.dims, 0, kl, lop.op.optype)
        rs, ra = S.fill_params(big.right.dims, 1, kr, rop.op.optype)
        fills[i] = (ls, la, rs, ra)
        materialise(big.left, lop.op, ls, la)
        materialise(big.right, rop.op, rs, ra)
        O.tensor_multiply(big, lop, rop, c, v, 0, scale)
    _, _, ref = refbench.run(big, terms, pick, cores=2, fills=fills, psi=psi)
    got = big.flatten(v)
    assert np.linalg.norm(got - ref) <= 1e-13 * np.linalg.norm(ref)


@needs_ref
@pytest.mark.gpu
def test_gpu_terms_equal_reference_on_synthetic_block():
    """CUDA TensorMultiply (C ABI) vs the real reference's on the benchmark's synthetic block (small M)."""
    a = args(300, left_sites=6, norbs=14, nelec=14)
    sb = S.make_big_block(norbs=a.norbs, nelec=a.nelec, M=a.M, left_sites=a.left_sites, device=0)
    try:
        psi = np.random.default_rng(4).standard_normal(sb.size)
        err, n = bench.reference_parity(sb, a, psi, nterms=20)
        assert n >= 10 and err < 1e-12
    finally:
        sb.close()


@pytest.mark.gpu
def test_gpu_synthetic_sigma_matches_oracle():
    """Whole multiplyH on a small synthetic block: CUDA path vs the numpy oracle with the device's operator values."""
    a = args(120)
    sb = S.make_big_block(norbs=a.norbs, nelec=a.nelec, M=a.M, left_sites=a.left_sites, device=0)
    try:
        big, terms = bench._oracle_big(a)
        for side, blk in enumerate((big.left, big.right)):
            for k, op in enumerate(blk.ops):
                data = sb.download_op(side, sb.op_ids[side][k])
                op.blocks.clear()
                off = 0
                for i, j in zip(*np.nonzero(op.allowed)):
                    n = int(blk.dims[i]) * int(blk.dims[j])
                    op.blocks[(int(i), int(j))] = data[off:off + n].reshape(int(blk.dims[i]), int(blk.dims[j]))
                    off += n
        psi = np.random.default_rng(5).standard_normal(sb.size)
        ref = big.flatten(O.multiply_h(big, big.unflatten(psi)))
        got = sb.multiplyH(psi)
        assert np.linalg.norm(got - ref) <= 1e-12 * np.linalg.norm(ref)
        for cls in (0, 1, 2):
            sb2 = S.make_big_block(norbs=a.norbs, nelec=a.nelec, M=a.M, left_sites=a.left_sites, device=0, options={"tile_class": cls, "multi_stream": 0})
            try:
                assert np.linalg.norm(sb2.multiplyH(psi) - ref) <= 1e-12 * np.linalg.norm(ref)
            finally:
                sb2.close()
    finally:
        sb.close()


@pytest.mark.gpu
def test_gpu_persistent_kernel_is_bit_identical():
    """The persistent 128 x 128 kernel (tiles claimed from an atomic counter, one pipeline across tile boundaries) against the
    one-CTA-per-tile kernel: same tiles, same K order inside a tile => bit-identical sigma; forced 128 x 128 tiles so that the
    class has far more tiles than SMs, and the default (auto) classes as well."""
    a = args(300, left_sites=6, norbs=14, nelec=14)
    psi = np.random.default_rng(11).standard_normal
    out = {}
    for cls in (0, -1):
        for pers in (0, 1):
            sb = S.make_big_block(norbs=a.norbs, nelec=a.nelec, M=a.M, left_sites=a.left_sites, device=0,
                                  options={"tile_class": cls, "persistent": pers})
            try:
                x = np.random.default_rng(11).standard_normal(sb.size)
                out[(cls, pers)] = sb.multiplyH(x)
                if cls == 0:
                    assert sb.plan_stats()["tiles"] > 148 * 4
            finally:
                sb.close()
    assert np.array_equal(out[(0, 0)], out[(0, 1)])
    assert np.array_equal(out[(-1, 0)], out[(-1, 1)])
    assert np.linalg.norm(out[(0, 1)] - out[(-1, 1)]) <= 1e-12 * np.linalg.norm(out[(-1, 1)])
