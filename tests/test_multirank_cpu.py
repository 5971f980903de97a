"""N > 1 host logic on CPU (gloo, world_size 2): the partition of multiplyH's operator terms across ranks
(processorindex / trimap_2d ownership, para_array.h:33-42,360-383) and the sum of the partial sigma vectors
(distributedaccumulate, distribute.h:42-76 -> all-reduce).  Each rank plans with the C library (planning-only context), computes
its partial sigma with the numpy oracle over ITS terms only, and the partial results are all-reduced over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, golden_path, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    from block_b200 import hotpath
    from oracle import dmrg_oracle as O
    from oracle import dumpio
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rec = dumpio.read_records(golden_path)
    big = dumpio.big_from(rec)
    sb = hotpath.spinblock_from_record(rec, device=-1, rank=rank, nranks=world)          # planning only: no GPU needed
    sb1 = hotpath.spinblock_from_record(rec, device=-1, rank=0, nranks=1)
    lo, ro, fl, sc, ow = sb.terms(all_ranks=False)
    alo, aro, afl, asc, aow = sb.terms(all_ranks=True)
    full = sb1.terms(all_ranks=True)
    # every rank sees the same global list, identical to the single-rank list except for the owner column
    assert np.array_equal(alo, full[0]) and np.array_equal(aro, full[1]) and np.array_equal(afl, full[2]) and np.array_equal(asc, full[3])
    mine = np.nonzero(aow == rank)[0]
    assert len(mine) == len(lo) and np.array_equal(alo[mine], lo) and np.array_equal(aro[mine], ro)
    # H_L x 1, 1 x H_R and e_core are rank 0's (spinblock.C:735-747)
    nfixed = 2 + (1 if abs(big.core_energy) > 1e-20 else 0)
    assert (aow[:nfixed] == 0).all()
    # the partition is exhaustive and disjoint; flop shares add up
    counts = torch.zeros(world, dtype=torch.int64); counts[rank] = len(lo)
    dist.all_reduce(counts)
    assert int(counts.sum()) == len(alo)
    fl_mine = torch.tensor([sb.sigma_flops(all_ranks=False)], dtype=torch.float64)
    dist.all_reduce(fl_mine)
    assert abs(fl_mine.item() - sb.sigma_flops(all_ranks=True)) <= 1e-9 * sb.sigma_flops(all_ranks=True)
    assert sb.sigma_flops(all_ranks=True) == sb1.sigma_flops(all_ranks=True)
    # partial sigma over this rank's terms (oracle arithmetic), then the all-reduce that replaces the reference's tree reduce
    terms = O.h_terms(big)
    c = big.unflatten(rec["rpsi"])
    v = big.zeros()
    for i in mine:
        lop, rop, scale = terms[int(i)]
        O.tensor_multiply(big, lop, rop, c, v, 0, scale)
    part = torch.from_numpy(big.flatten(v))
    dist.all_reduce(part)
    err = np.linalg.norm(part.numpy() - rec["rsigma"]) / np.linalg.norm(rec["rsigma"])
    assert err < 1e-12, err
    # ownership follows the reference's index maps
    for i in mine[nfixed if rank == 0 else 0:]:
        lop, rop, _ = terms[int(i)]
        loop_op = lop.op if big.left.loop and len(lop.op.orbs) == 2 else (rop.op if len(rop.op.orbs) == 2 and not big.left.loop else None)
        cre = next((o for o in (lop.op, rop.op) if o.optype == O.CRE), None)
        if cre is not None:
            assert cre.orbs[0] % world == rank
    sb.close(); sb1.close()
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    dist.destroy_process_group()


@pytest.mark.parametrize("which", [0, -1])
def test_term_partition_and_allreduce_world2(tmp_path, which):
    from conftest import GOLDEN
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), GOLDEN[which], str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / ("ok%d" % r)) for r in range(world))
