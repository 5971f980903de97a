"""SURVEY.md N3 at the C ABI: the renormalised block stays on the device between block iterations (b2d_cache_put_rotated / b2d_cache_use),
moves to pinned host memory under memory pressure (b2d_cache_spill = what every failing device allocation of the library does by itself)
and comes back bit for bit - the role of SpinBlock::store / restore (save_load_block.C:23-108) and of the reference's scratch disk."""
import numpy as np
import pytest

from block_b200 import hotpath
from oracle import dumpio

pytestmark = pytest.mark.gpu


def rotated_block(golden):
    rec, big = golden
    sb = hotpath.spinblock_from_record(rec, device=0)
    sb.set_rotation_matrices(dumpio.rotation_from(rec))
    old, dims, ops = sb.transform_operators()
    return rec, sb, dims, ops


def test_cached_block_is_the_rotated_block_and_survives_a_spill(golden):
    rec, sb, dims, ops = rotated_block(golden)
    try:
        ids = list(sb.op_ids[0])
        tok = sb.cache_put_rotated()
        nq, nops, _ = sb.cache_block_info(tok)
        assert nq == len(dims) and nops == len(ops)
        st = sb.cache_stats()
        assert st["entries"] == 1 and st["device_doubles"] > 0 and st["spilled_doubles"] == 0
        for k, (allowed, data) in enumerate(ops):
            a2, d2 = sb.cache_download_op(tok, ids[k])
            assert np.array_equal(a2, allowed) and np.array_equal(d2, data)          # bit for bit
        sb.release_block()                                   # the next block iteration: nothing is a child any more
        sb.cache_spill(1e15)                                 # more than the GPU has: every entry that may move does
        st = sb.cache_stats()
        assert st["device_doubles"] == 0 and st["spilled_doubles"] > 0 and st["evictions"] == 1, st
        for k, (allowed, data) in enumerate(ops):
            a2, d2 = sb.cache_download_op(tok, ids[k])
            assert np.array_equal(a2, allowed) and np.array_equal(d2, data)
        # a spilled entry becomes a child again (one H2D copy into the arena): the operators on the device are the same bits
        sb.cache_use(tok, 0)
        for k in (0, len(ops) // 2, len(ops) - 1):
            got = sb.download_op(0, ids[k])
            assert np.array_equal(got, ops[k][1]), k
        sb.cache_drop(tok)
        assert sb.cache_stats()["entries"] == 0
    finally:
        sb.close()


def test_children_of_the_current_block_iteration_are_not_evicted(golden):
    rec, sb, dims, ops = rotated_block(golden)
    try:
        sb_ids = list(sb.op_ids[0])
        tok = sb.cache_put_rotated()
        sb.release_block()
        sb.cache_use(tok, 0)                                 # in place: the child's operators point into the cached buffer
        sb.cache_spill(1e15)
        st = sb.cache_stats()
        assert st["evictions"] == 0 and st["device_doubles"] > 0, st
        assert np.array_equal(sb.download_op(0, sb_ids[0]), ops[0][1])
        sb.release_block()
        sb.cache_spill(1e15)
        assert sb.cache_stats()["evictions"] == 1
    finally:
        sb.close()
