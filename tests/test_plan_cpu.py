"""Host-side (integer / scalar) logic of the product against the oracle, on a planning-only context: no GPU needed.
Covers: C-ABI export list, psi layout (bit-exact), term list of multiplyH (operators, transposes, scale factors),
algorithmic flop count, term ownership across ranks, and loud failure of compute calls without a device."""
import ctypes as C
import re
import os

import numpy as np
import pytest

from block_b200 import _lib, hotpath
from oracle import dmrg_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def planner(rec, rank=0, nranks=1):
    return hotpath.spinblock_from_record(rec, device=-1, rank=rank, nranks=nranks)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "block_b200.h")).read()
    declared = set(re.findall(r"\b(b2d_[a-zA-Z0-9_]+)\s*\(", header))
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.b2d_abi_version() == 4


def test_compute_without_device_fails_loudly(golden):
    rec, _ = golden
    sb = planner(rec)
    with pytest.raises(hotpath.B2DError, match="no CPU fallback"):
        sb.multiplyH(np.zeros(sb.size))
    with pytest.raises(hotpath.B2DError):
        sb.diagonalH()


def test_psi_layout_bit_exact(golden):
    rec, big = golden
    sb = planner(rec)
    assert sb.size == rec["rpsi"].size
    l, r, off = sb.psi_blocks()
    assert list(l) == list(rec["big.lmap"]) and list(r) == list(rec["big.rmap"])
    assert list(off) == list(rec["big.unblocked"])


def test_term_list_matches_oracle(golden):
    rec, big = golden
    sb = planner(rec)
    lo, ro, fl, sc, ow = sb.terms(all_ranks=True)
    ref = O.h_terms(big)
    assert len(ref) == len(lo)
    for k, (lv, rv, scale) in enumerate(ref):
        assert big.left.ops[lo[k]] is lv.op and big.right.ops[ro[k]] is rv.op
        assert bool(fl[k] & 1) == lv.t and bool(fl[k] & 2) == rv.t
        assert sc[k] == pytest.approx(scale, rel=1e-14, abs=0)
    assert (ow == 0).all()


def test_flops_match_oracle(golden):
    rec, big = golden
    sb = planner(rec)
    assert sb.sigma_flops() == O.sigma_flops(big)
    st = sb.plan_stats()
    assert 0 < st["flops_executed"] <= sb.sigma_flops()


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_term_partition_follows_reference_rule(golden, nranks):
    rec, big = golden
    norbs = len(rec["spin_orbs_symmetry"]) // 2
    seen = []
    total = None
    for rank in range(nranks):
        sb = planner(rec, rank, nranks)
        lo, ro, fl, sc, ow = sb.terms(all_ranks=True)
        total = len(lo)
        mine = sb.terms(all_ranks=False)
        idx = [k for k in range(total) if ow[k] == rank]
        assert list(mine[0]) == [lo[k] for k in idx] and list(mine[1]) == [ro[k] for k in idx]
        seen += idx
        # ownership rule: processorindex(i) = i % size (para_array.h:33-42); trimap_2d(i,j,length) % size (:360-383)
        for k in range(total):
            lop, rop = big.left.ops[lo[k]], big.right.ops[ro[k]]
            orbs = lop.orbs or rop.orbs
            if len(orbs) == 0:
                assert ow[k] == 0
            elif len(orbs) == 1:
                assert ow[k] == orbs[0] % nranks
            else:
                i, j = max(orbs), min(orbs)
                half = norbs // 2
                tri = lambda x: x * (x + 1) // 2
                if i >= half and j >= half:
                    t = tri(norbs - j - 1) + norbs - i - 1
                elif i < half and j < half:
                    t = tri(norbs - half - 1) + norbs - half + tri(i) + j
                else:
                    t = tri(norbs - half - 1) + norbs - half + tri(half) + (i - half) * half + j
                assert ow[k] == t % nranks
    assert sorted(seen) == list(range(total))      # every term executed exactly once


def test_cost_weighted_term_ownership_balances_flops():
    """Option balance_terms ("allow a cost-weighted reassignment as long as the sum is unchanged"): the union of the ranks'
    term lists is the single-rank list, the sets are disjoint, and the executed flops per rank are within 1 % of the mean where the
    reference's static rule (i % n, trimap_2d % n) is several per cent off.  Planning only (no device)."""
    from block_b200 import synthetic as S
    world = 8
    shares = {}
    for balance in (0, 1):
        seen, fl = [], []
        for rank in range(world):
            sb = S.make_big_block(norbs=24, nelec=24, M=300, left_sites=11, device=-1, rank=rank, nranks=world, options={"balance_terms": balance}, fill=False)
            lo, ro, flg, sc, ow = sb.terms(all_ranks=True)
            mine = np.nonzero(ow == rank)[0]
            mlo, mro, mfl, msc, _ = sb.terms(all_ranks=False)
            assert np.array_equal(lo[mine], mlo) and np.array_equal(ro[mine], mro) and np.array_equal(flg[mine], mfl)
            seen.append(set(int(i) for i in mine))
            fl.append(sb.plan_stats()["flops_executed"])
            total_terms = len(lo)
            sb.close()
        assert sum(len(x) for x in seen) == total_terms and len(set().union(*seen)) == total_terms     # exhaustive and disjoint
        shares[balance] = max(fl) / (sum(fl) / world)
    assert shares[1] < 1.01, shares
    assert shares[1] <= shares[0]
    print("max / mean executed flops per rank at %d ranks: reference rule %.4f, cost-weighted %.4f" % (world, shares[0], shares[1]))


def test_narrow_tiles_get_their_own_split_k_family():
    """Option slice_iters_narrow: the narrow tiles of a sigma block (remainder bands of ragged sectors) are cut into more, shorter split-K
    slices than its 128 x 128 tiles - a second family of slice groups restricted to the narrow tile classes.  The work is unchanged (same
    executed flops, same useful flops inside the tiles), only the number of tiles grows.  Planning only (no device)."""
    from block_b200 import synthetic as S
    stats = {}
    for narrow in (0, 32):
        sb = S.make_big_block(norbs=24, nelec=24, M=600, left_sites=11, device=-1, options={"slice_iters": 64, "slice_iters_narrow": narrow}, fill=False)
        stats[narrow] = sb.plan_stats()
        sb.close()
    a, b = stats[0], stats[32]
    assert a["flops_executed"] == b["flops_executed"]
    assert abs(a["flops_in_tiles"] - b["flops_in_tiles"]) <= 1e-9 * a["flops_in_tiles"]
    assert b["tiles"] > a["tiles"], (a["tiles"], b["tiles"])
    assert a["step1_contractions"] == b["step1_contractions"]
