"""Host-side mirror of the reference's entry points for the sweep hot path, on top of the C ABI.

Names follow sanshar/Block (file:line under the reference root):
    SpinBlock.multiplyH            spinblock.h:235   spinblock.C:722
    SpinBlock.diagonalH            spinblock.h:240   spinblock.C:855
    SpinBlock.RenormaliseFrom      spinblock.h:247   renormalise.C:39
    SpinBlock.transform_operators  spinblock.h:253   save_load_block.C:267
    block_davidson                 linear.h:28       linear.C:179
    TensorMultiply                 operatorfunctions.h:46   operatorfunctions.C:485
Wavefunctions cross this boundary as flat float64 arrays in Wavefunction::FlattenInto order (wavefunction.C:167);
operators as their allowed sector blocks, row-major, (i outer, j inner).

Python is used here only because the tests and bench.py are Python; the C++ mirror with the reference's exact
signatures is block_b200/host/ (see INTEGRATION.md).  All arithmetic happens in libblockb200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib

HUBBARD = 1  # hamTypes, input.h


class B2DError(RuntimeError):
    pass


def _p(a, t):
    return a.ctypes.data_as(t)


@dataclass
class OperatorSpec:
    """One SparseMatrix of an operator array (BaseOperator.h:75-243)."""
    optype: int
    orbs: tuple
    comp: int
    dq: tuple
    fermion: bool
    allowed: np.ndarray            # (nq, nq) bool / uint8
    data: np.ndarray | None        # packed blocks or None (device-allocated zeros, e.g. for synthetic fills)


@dataclass
class BlockSpec:
    """StateInfo + operators of one child of the big block."""
    q: np.ndarray                  # (nq, 3) int: N, 2S, irrep
    dims: np.ndarray
    sites: tuple = ()
    loop: bool = False
    ops: list = field(default_factory=list)


class SpinBlock:
    """The `big` block of one block iteration (left child x right child) resident on one GPU."""

    def __init__(self, left: BlockSpec, right: BlockSpec, psi_dq, core_energy=0.0, hubbard=False, norbs=None,
                 device=0, rank=0, nranks=1, options=None):
        self.lib = _lib.load()
        self._ctx = C.c_void_p()
        rc = self.lib.b2d_create(int(device), C.byref(self._ctx))
        if rc:
            raise B2DError("b2d_create: " + self.lib.b2d_last_error(None).decode())
        self.device = device
        for k, v in (options or {}).items():
            self._ck(self.lib.b2d_set_option(self._ctx, k.encode(), float(v)))
        self._describe(left, right, psi_dq, core_energy, hubbard, norbs, rank, nranks)

    @classmethod
    def from_products(cls, left_parts, right_parts, psi_dq, core_energy=0.0, hubbard=False, norbs=None, device=0, rank=0, nranks=1, options=None,
                      integrals=None, fill_seed=20260):
        """The big block of a two-dot step from the blocks a sweep actually holds: each child is (renormalised block) x (one-site dot),
        given as (child BlockSpec, dot BlockSpec, product tables, BlockSpec of the enlarged block: sectors + operator list).  Every
        operator of the enlarged blocks is built on the device from the grandchildren (b2d_build_enlarged_op) - as factor lists with
        option "factorised", by the scatter kernel otherwise -, the blocks are parked (b2d_stash_product) and assembled.  Grandchild
        operators without data are allocated and filled with the counter-based random stream (synthetic benchmark)."""
        self = cls.__new__(cls)
        self.lib = _lib.load()
        self._ctx = C.c_void_p()
        if self.lib.b2d_create(int(device), C.byref(self._ctx)):
            raise B2DError("b2d_create: " + self.lib.b2d_last_error(None).decode())
        self.device = device
        for k, v in (options or {}).items():
            self._ck(self.lib.b2d_set_option(self._ctx, k.encode(), float(v)))
        if integrals is not None:
            v1, v2, irr = (np.ascontiguousarray(integrals[0], np.float64), np.ascontiguousarray(integrals[1], np.float64), np.ascontiguousarray(integrals[2], np.int32))
            self._ck(self.lib.b2d_set_integrals(self._ctx, len(irr), _p(v1, _lib.c_f64p), _p(v2, _lib.c_f64p), _p(irr, _lib.c_i32p), 1e-15, 1e-15))
        self.op_ids = [[], []]
        enlarged = []
        for slot, (child, dot, tables, big) in enumerate((left_parts, right_parts)):
            for side, blk in enumerate((child, dot)):
                q = np.ascontiguousarray(blk.q, dtype=np.int32).reshape(-1, 3)
                dims = np.ascontiguousarray(blk.dims, dtype=np.int32)
                sites = np.ascontiguousarray(blk.sites, dtype=np.int32)
                self._ck(self.lib.b2d_set_block(self._ctx, side, len(dims), _p(q, _lib.c_i32p), _p(dims, _lib.c_i32p), int(blk.loop), len(sites), _p(sites, _lib.c_i32p)))
                ids = [self._add_op(side, op) for op in blk.ops]
                if device >= 0 and any(op.data is None for op in blk.ops):
                    self._ck(self.lib.b2d_alloc_ops(self._ctx, side))
                    amp = 1.0 / float(np.sqrt(np.sum(dims)))
                    for k, op in enumerate(blk.ops):
                        if op.data is None:
                            self._ck(self.lib.b2d_fill_op_random(self._ctx, side, ids[k], int(fill_seed) * 1000003 + slot * 500009 + side * 250007 + k, amp, 0))
            tq = np.ascontiguousarray(tables["q"], np.int32).reshape(-1, 3)
            td = np.ascontiguousarray(tables["dims"], np.int32)
            lm, rm, ud = (np.ascontiguousarray(tables[k], np.int32) for k in ("unc.lmap", "unc.rmap", "unc.dims"))
            ob, on = np.ascontiguousarray(tables["old_to_new_begin"], np.int32), np.ascontiguousarray(tables["old_to_new"], np.int32)
            self._ck(self.lib.b2d_set_product_stateinfo(self._ctx, len(td), _p(tq, _lib.c_i32p), _p(td, _lib.c_i32p), len(ud), _p(lm, _lib.c_i32p),
                                                         _p(rm, _lib.c_i32p), _p(ud, _lib.c_i32p), _p(ob, _lib.c_i32p), _p(on, _lib.c_i32p)))
            assert np.array_equal(tq, np.asarray(big.q, np.int32).reshape(-1, 3)) and np.array_equal(td, np.asarray(big.dims, np.int32))
            for op in big.ops:
                o = np.asarray(list(op.orbs) + [-1, -1], dtype=np.int32)
                dq = np.asarray(op.dq, dtype=np.int32)
                pid = C.c_int(-1)
                self._ck(self.lib.b2d_build_enlarged_op(self._ctx, int(op.optype), len(op.orbs), _p(o, _lib.c_i32p), int(op.comp), _p(dq, _lib.c_i32p),
                                                        int(bool(op.fermion)), int(bool(hubbard)), C.byref(pid)))
                self.op_ids[slot].append(pid.value)
            sites = np.ascontiguousarray(big.sites, dtype=np.int32)
            self._ck(self.lib.b2d_stash_product(self._ctx, slot, int(big.loop), len(sites), _p(sites, _lib.c_i32p)))
            enlarged.append(big)
        self._ck(self.lib.b2d_assemble_big(self._ctx))
        self.left, self.right = enlarged
        if norbs is None:
            norbs = len(self.left.sites) + len(self.right.sites)
        dq = np.asarray(psi_dq, dtype=np.int32)
        self._ck(self.lib.b2d_plan(self._ctx, _p(dq, _lib.c_i32p), float(core_energy), int(bool(hubbard)), int(norbs), int(rank), int(nranks)))
        self.size = int(self.lib.b2d_psi_size(self._ctx))
        self.rank, self.nranks = rank, nranks
        self._nslots = 0
        return self

    def reset(self, left: BlockSpec, right: BlockSpec, psi_dq, core_energy=0.0, hubbard=False, norbs=None, rank=0, nranks=1):
        """b2d_reset + a new block description on the SAME context (streams, arena slabs and scratch buffers are kept): what a
        sweep does between block iterations."""
        self._ck(self.lib.b2d_reset(self._ctx))
        self._describe(left, right, psi_dq, core_energy, hubbard, norbs, rank, nranks)

    def _describe(self, left, right, psi_dq, core_energy, hubbard, norbs, rank, nranks):
        self.left, self.right = left, right
        self.op_ids = [[], []]
        for side, blk in enumerate((left, right)):
            q = np.ascontiguousarray(blk.q, dtype=np.int32).reshape(-1, 3)
            dims = np.ascontiguousarray(blk.dims, dtype=np.int32)
            sites = np.ascontiguousarray(blk.sites, dtype=np.int32)
            self._ck(self.lib.b2d_set_block(self._ctx, side, len(dims), _p(q, _lib.c_i32p), _p(dims, _lib.c_i32p), int(blk.loop),
                                            len(sites), _p(sites, _lib.c_i32p)))
            for op in blk.ops:
                self.op_ids[side].append(self._add_op(side, op))
        if norbs is None:
            norbs = len(left.sites) + len(right.sites)
        dq = np.asarray(psi_dq, dtype=np.int32)
        self._ck(self.lib.b2d_plan(self._ctx, _p(dq, _lib.c_i32p), float(core_energy), int(bool(hubbard)), int(norbs), int(rank), int(nranks)))
        self.size = int(self.lib.b2d_psi_size(self._ctx))
        self.rank, self.nranks = rank, nranks
        self._nslots = 0

    # -- plumbing ---------------------------------------------------------------------------------------------
    def _ck(self, rc):
        if rc:
            raise B2DError("[%d] %s" % (rc, self.lib.b2d_last_error(self._ctx).decode()))

    def _add_op(self, side, op: OperatorSpec):
        orbs = np.asarray(list(op.orbs) + [-1, -1], dtype=np.int32)
        dq = np.asarray(op.dq, dtype=np.int32)
        allowed = np.ascontiguousarray(op.allowed, dtype=np.uint8)
        data = None if op.data is None else np.ascontiguousarray(op.data, dtype=np.float64)
        oid = C.c_int(-1)
        self._ck(self.lib.b2d_add_op(self._ctx, side, int(op.optype), len(op.orbs), _p(orbs, _lib.c_i32p), int(op.comp), _p(dq, _lib.c_i32p),
                                     int(bool(op.fermion)), _p(allowed, _lib.c_u8p), None if data is None else _p(data, _lib.c_f64p), C.byref(oid)))
        return oid.value

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self.lib.b2d_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reserve(self, n):
        if n > self._nslots:
            self._ck(self.lib.b2d_vec_reserve(self._ctx, n))
            self._nslots = n

    def upload(self, slot, flat):
        self.reserve(slot + 1)
        flat = np.ascontiguousarray(flat, dtype=np.float64)
        assert flat.size == self.size
        self._ck(self.lib.b2d_vec_upload(self._ctx, slot, _p(flat, _lib.c_f64p)))

    def download(self, slot):
        out = np.empty(self.size)
        self._ck(self.lib.b2d_vec_download(self._ctx, slot, _p(out, _lib.c_f64p)))
        return out

    def set_option(self, key, value):
        self._ck(self.lib.b2d_set_option(self._ctx, key.encode(), float(value)))

    def attach_communicator(self, unique_id: bytes, rank, nranks):
        import os
        path = _lib.nccl_library_path()
        if path and "B2D_NCCL_LIB" not in os.environ:
            os.environ["B2D_NCCL_LIB"] = path
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._ck(self.lib.b2d_comm_init(self._ctx, buf, rank, nranks))

    @staticmethod
    def nccl_unique_id() -> bytes:
        import os
        lib = _lib.load()
        path = _lib.nccl_library_path()
        if path and "B2D_NCCL_LIB" not in os.environ:
            os.environ["B2D_NCCL_LIB"] = path
        buf = (C.c_uint8 * 128)()
        if lib.b2d_nccl_unique_id(buf):
            raise B2DError(lib.b2d_last_error(None).decode())
        return bytes(buf)

    # -- integer side: layout and term list -------------------------------------------------------------------
    def psi_blocks(self):
        n = self.lib.b2d_psi_num_blocks(self._ctx)
        l, r, o = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.int64)
        self._ck(self.lib.b2d_psi_blocks(self._ctx, _p(l, _lib.c_i32p), _p(r, _lib.c_i32p), _p(o, _lib.c_i64p)))
        return l, r, o

    def terms(self, all_ranks=False):
        n = self.lib.b2d_num_terms(self._ctx, int(all_ranks))
        lo, ro, fl, ow = (np.empty(n, np.int32) for _ in range(4))
        sc = np.empty(n)
        self._ck(self.lib.b2d_terms(self._ctx, int(all_ranks), _p(lo, _lib.c_i32p), _p(ro, _lib.c_i32p), _p(fl, _lib.c_i32p), _p(sc, _lib.c_f64p),
                                    _p(ow, _lib.c_i32p)))
        return lo, ro, fl, sc, ow

    def sigma_flops(self, all_ranks=True):
        return float(self.lib.b2d_sigma_flops(self._ctx, int(all_ranks)))

    def plan_stats(self):
        out = np.zeros(13)
        self._ck(self.lib.b2d_plan_stats(self._ctx, _p(out, _lib.c_f64p), 13))
        keys = ["chunks", "step1_contractions", "step2_segments", "tiles", "workspace_doubles", "arena_doubles", "launches_per_sigma", "flops_executed",
                "flops_in_tiles", "flops_issued", "combo_doubles", "factors_direct", "factors_combo"]
        return dict(zip(keys, out.tolist()))

    # -- the reference's entry points ---------------------------------------------------------------------------
    def multiplyH(self, c, v=None):
        """SpinBlock::multiplyH(Wavefunction& c, Wavefunction* v, int): v += H c on host buffers (returns v)."""
        c = np.ascontiguousarray(c, dtype=np.float64)
        acc = v is not None
        if v is None:
            v = np.zeros(self.size)
        assert c.size == self.size and v.size == self.size and v.flags.c_contiguous
        self._ck(self.lib.b2d_multiplyH_host(self._ctx, _p(c, _lib.c_f64p), _p(v, _lib.c_f64p), int(acc)))
        return v

    def multiplyH_into(self, c, v, accumulate=False):
        """multiplyH with caller-owned host buffers (pinned buffers make the copies true DMA): v (+)= H c."""
        assert c.size == self.size and v.size == self.size and c.flags.c_contiguous and v.flags.c_contiguous
        self._ck(self.lib.b2d_multiplyH_host(self._ctx, _p(c, _lib.c_f64p), _p(v, _lib.c_f64p), int(accumulate)))

    def sigma(self, src_slot, dst_slot, accumulate=False):
        """multiplyH on device-resident slots (what block_davidson uses internally)."""
        self._ck(self.lib.b2d_sigma(self._ctx, src_slot, dst_slot, int(accumulate)))

    def TensorMultiply(self, left_op, right_op, c, v, left_transposed=False, right_transposed=False, opq_spin=0, scale=1.0):
        """operatorfunctions::TensorMultiply(ablock, a, b, cblock, c, v, opQ, scale): v += scale (a x b) c."""
        self.upload(0, c)
        self.upload(1, v)
        flags = (1 if left_transposed else 0) | (2 if right_transposed else 0)
        self._ck(self.lib.b2d_tensor_multiply(self._ctx, left_op, right_op, flags, int(opq_spin), float(scale), 0, 1))
        return self.download(1)

    def tensor_multiply_slots(self, left_op, right_op, left_transposed, right_transposed, opq_spin, scale, src_slot, dst_slot):
        """TensorMultiply on device-resident slots: dst += scale (a x b) src."""
        flags = (1 if left_transposed else 0) | (2 if right_transposed else 0)
        self._ck(self.lib.b2d_tensor_multiply(self._ctx, left_op, right_op, flags, int(opq_spin), float(scale), src_slot, dst_slot))

    def clear(self, slot):
        self.reserve(slot + 1)
        self._ck(self.lib.b2d_vec_clear(self._ctx, slot))

    def diagonalH(self, slot=None):
        """SpinBlock::diagonalH(DiagonalMatrix&): diag(H) in flat psi order."""
        s = 0 if slot is None else slot
        self.reserve(s + 1)
        self._ck(self.lib.b2d_diagonal(self._ctx, s))
        return self.download(s)

    def block_davidson(self, guesses, diag, normtol, deflation_min=2, deflation_max=20, lowerStates=()):
        """Linear::block_davidson(b, h_diag, normtol, warmUp, h_multiply, useprecond, currentRoot, lowerStates): returns
        (eigenvalues, solutions, number of H applications).  lowerStates non-empty = the state-specific form."""
        n, nl = len(guesses), len(lowerStates)
        self.reserve(n + 1 + nl)
        for i, g in enumerate(guesses):
            self.upload(i, g)
        self.upload(n, diag)
        for i, l in enumerate(lowerStates):
            self.upload(n + 1 + i, l)
        ev = np.zeros(n)
        nm = C.c_int(0)
        res = C.c_double(0.0)
        self._ck(self.lib.b2d_davidson_lower(self._ctx, n, 0, n, float(normtol), int(deflation_min), int(deflation_max), nl, n + 1,
                                             _p(ev, _lib.c_f64p), C.byref(nm), C.byref(res)))
        return ev, [self.download(i) for i in range(n)], nm.value

    def make_density(self, waves, weights, noise=0.0):
        """DensityMatrix::makedensitymatrix(wave_solutions, big, weights, noise, 0, warmup): returns the per-sector blocks."""
        n = len(waves)
        self.reserve(n)
        for i, w in enumerate(waves):
            self.upload(i, w)
        return self._density_from_slots(n, weights, noise)

    def _density_from_slots(self, n, weights, noise=0.0):
        w = np.ascontiguousarray(weights, dtype=np.float64)
        self._ck(self.lib.b2d_make_density(self._ctx, n, 0, _p(w, _lib.c_f64p)))
        if noise > 0.0:
            self._ck(self.lib.b2d_add_onedot_noise(self._ctx, n, 0, float(noise)))
        return self.density()

    def add_wavefunction_density(self, dq, flat, weight):
        """rho += weight * w w^T for a host wavefunction of target quantum dq (the device half of add_twodot_noise, density.C:92-165)."""
        q = np.asarray(dq, dtype=np.int32)
        flat = np.ascontiguousarray(flat, dtype=np.float64)
        assert flat.size == self.wavefunction_size(dq)
        self._ck(self.lib.b2d_add_wavefunction_density(self._ctx, _p(q, _lib.c_i32p), _p(flat, _lib.c_f64p), float(weight)))

    def wavefunction_size(self, dq):
        q = np.asarray(dq, dtype=np.int32)
        return int(self.lib.b2d_wavefunction_size(self._ctx, _p(q, _lib.c_i32p)))

    def TensorMultiplyOne(self, side, op_id, c, v, dst_dq, transposed=False, scale=1.0):
        """One-operator operatorfunctions::TensorMultiply(ablock, a, cblock, c, v, dQ, scale): v += scale (a x 1) c or (1 x a) c;
        v is a flat wavefunction with target quantum dst_dq (returns v)."""
        c = np.ascontiguousarray(c, dtype=np.float64)
        q = np.asarray(dst_dq, dtype=np.int32)
        assert c.size == self.size and v.size == self.wavefunction_size(dst_dq) and v.flags.c_contiguous
        self._ck(self.lib.b2d_tensor_multiply_one_host(self._ctx, int(side), int(op_id), int(bool(transposed)), _p(q, _lib.c_i32p), float(scale),
                                                       _p(c, _lib.c_f64p), _p(v, _lib.c_f64p)))
        return v

    def density(self):
        flat = np.empty(int(self.lib.b2d_density_size(self._ctx)))
        self._ck(self.lib.b2d_density_download(self._ctx, _p(flat, _lib.c_f64p)))
        out, off = [], 0
        for d in self.left.dims:
            d = int(d)
            out.append(flat[off:off + d * d].reshape(d, d)); off += d * d
        return out

    def set_density(self, blocks):
        """Upload a density matrix (one d_q x d_q block per left sector): what a binding does when the caller hands diagonalise_dm a matrix."""
        flat = np.concatenate([np.ascontiguousarray(b, dtype=np.float64).ravel() for b in blocks])
        assert flat.size == int(self.lib.b2d_density_size(self._ctx))
        self._ck(self.lib.b2d_density_upload(self._ctx, _p(flat, _lib.c_f64p)))

    def diagonalise_dm(self):
        """diagonalise_dm (rotationmat.C:258): per-sector eigenvalues (ascending, < 1e-14 -> 0)."""
        ev = np.empty(int(np.sum(self.left.dims)))
        self._ck(self.lib.b2d_diagonalise_dm(self._ctx, _p(ev, _lib.c_f64p)))
        out, off = [], 0
        for d in self.left.dims:
            out.append(ev[off:off + int(d)]); off += int(d)
        return out

    def select_states(self, keep_states):
        """sort_weights + assign_matrix_by_dm: returns (kept counts per sector, discarded weight, rotation matrices)."""
        kept = np.zeros(len(self.left.dims), np.int32)
        err = C.c_double(0.0)
        self._ck(self.lib.b2d_select_states(self._ctx, int(keep_states), _p(kept, _lib.c_i32p), C.byref(err)))
        return kept, err.value, self.rotation_matrices(kept)

    def rotation_matrices(self, kept):
        n = int(self.lib.b2d_rotation_size(self._ctx))
        flat = np.empty(max(n, 1))
        self._ck(self.lib.b2d_rotation_download(self._ctx, _p(flat, _lib.c_f64p)))
        out, off = [], 0
        for d, k in zip(self.left.dims, kept):
            d, k = int(d), int(k)
            out.append(flat[off:off + d * k].reshape(d, k).copy()); off += d * k
        return out

    def set_rotation_matrices(self, rot):
        kept = np.asarray([r.shape[1] for r in rot], dtype=np.int32)
        flat = np.concatenate([np.ascontiguousarray(r, dtype=np.float64).ravel() for r in rot] + [np.zeros(1)])
        self._ck(self.lib.b2d_rotation_upload(self._ctx, _p(kept, _lib.c_i32p), _p(flat, _lib.c_f64p)))

    def RenormaliseFrom(self, guesses, weights, normtol, keep_states, deflation_min=2, deflation_max=20, noise=0.0):
        """SpinBlock::RenormaliseFrom (two-dot): returns dict(energies, kept, error, n_multiply, rotation)."""
        n = len(guesses)
        self.reserve(n + 1)
        for i, g in enumerate(guesses):
            self.upload(i, g)
        w = np.ascontiguousarray(weights, dtype=np.float64)
        en = np.zeros(n)
        kept = np.zeros(len(self.left.dims), np.int32)
        err = C.c_double(0.0)
        nm = C.c_int(0)
        self._ck(self.lib.b2d_renormalise_from(self._ctx, n, 0, _p(w, _lib.c_f64p), float(normtol), int(keep_states), int(deflation_min),
                                               int(deflation_max), float(noise), _p(en, _lib.c_f64p), _p(kept, _lib.c_i32p), C.byref(err), C.byref(nm)))
        return dict(energies=en, kept=kept, error=err.value, n_multiply=nm.value, rotation=self.rotation_matrices(kept),
                    solutions=[self.download(i) for i in range(n)])

    def transform_operators(self):
        """SpinBlock::transform_operators(rotateMatrix) on the LEFT child: returns (old sector index per new sector,
        new dims, [(allowed, packed data)] per left operator in registration order)."""
        self._ck(self.lib.b2d_transform_operators(self._ctx))
        nq = self.lib.b2d_rotated_num_sectors(self._ctx)
        old, dims = np.zeros(nq, np.int32), np.zeros(nq, np.int32)
        self._ck(self.lib.b2d_rotated_sectors(self._ctx, _p(old, _lib.c_i32p), _p(dims, _lib.c_i32p)))
        ops = []
        for oid in self.op_ids[0]:
            n = int(self.lib.b2d_rotated_op_size(self._ctx, oid))
            allowed = np.zeros((nq, nq), np.uint8)
            data = np.empty(max(n, 1))
            self._ck(self.lib.b2d_rotated_op_download(self._ctx, oid, _p(allowed, _lib.c_u8p), _p(data, _lib.c_f64p)))
            ops.append((allowed.astype(bool), data[:n]))
        return old, dims, ops

    # -- device-side shadow of the scratch files (save_load_block.C:23-108) ----------------------------------
    def cache_put_rotated(self):
        """The renormalised left block that transform_operators left on the device becomes a cache entry: returns its token."""
        tok = C.c_uint64(0)
        self._ck(self.lib.b2d_cache_put_rotated(self._ctx, C.byref(tok)))
        return int(tok.value)

    def cache_block_info(self, token):
        nq, nops, ns = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        self._ck(self.lib.b2d_cache_block_info(self._ctx, C.c_uint64(token), C.byref(nq), C.byref(nops), C.byref(ns)))
        return nq.value, nops.value, ns.value

    def cache_download_op(self, token, op_id):
        nq, _, _ = self.cache_block_info(token)
        size = C.c_int64(0)
        self._ck(self.lib.b2d_cache_op_info(self._ctx, C.c_uint64(token), op_id, None, None, None, None, C.byref(size)))
        allowed = np.zeros((nq, nq), np.uint8)
        data = np.empty(max(int(size.value), 1))
        self._ck(self.lib.b2d_cache_download_op(self._ctx, C.c_uint64(token), op_id, _p(allowed, _lib.c_u8p), _p(data, _lib.c_f64p)))
        return allowed.astype(bool), data[:int(size.value)]

    def cache_use(self, token, side, is_loop=False):
        self._ck(self.lib.b2d_cache_use(self._ctx, C.c_uint64(token), int(side), int(is_loop)))

    def cache_drop(self, token):
        self._ck(self.lib.b2d_cache_drop(self._ctx, C.c_uint64(token)))

    def cache_spill(self, nbytes):
        self._ck(self.lib.b2d_cache_spill(self._ctx, float(nbytes)))

    def cache_stats(self):
        out = np.zeros(6)
        self._ck(self.lib.b2d_cache_stats(self._ctx, _p(out, _lib.c_f64p), 6))
        return dict(zip(["entries", "device_doubles", "spilled_doubles", "puts", "uses", "evictions"], out))

    def release_block(self):
        """b2d_reset: forget the block description (the cache entries stay)."""
        self._ck(self.lib.b2d_reset(self._ctx))

    # -- measurement ---------------------------------------------------------------------------------------------
    def last_timing_ms(self):
        out = np.zeros(4)
        self._ck(self.lib.b2d_last_timing(self._ctx, _p(out, _lib.c_f64p), 4))
        return out

    def sigma_profile(self, src_slot, dst_slot):
        """One multiplyH with events around every contraction launch: dict[(step, tile_class)] -> (ms, useful flops,
        issued flops, launches); tile class 9 = the warp-per-block kernel of the tiny (<= 8 x 8) sectors."""
        out = np.zeros(80)
        self._ck(self.lib.b2d_sigma_profile(self._ctx, src_slot, dst_slot, _p(out, _lib.c_f64p)))
        return {(st, c): tuple(out[(st * 10 + c) * 4:(st * 10 + c) * 4 + 4]) for st in range(2) for c in range(10)}

    def kernel_launches(self):
        return int(self.lib.b2d_kernel_launches(self._ctx))

    def sync(self):
        self._ck(self.lib.b2d_sync(self._ctx))

    def measure_fp64_peak(self):
        a, b = C.c_double(0), C.c_double(0)
        self._ck(self.lib.b2d_measure_fp64_peak(self._ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def measure_level1(self, reps=10):
        """HBM-bound Davidson kernels timed alone: dict name -> (ms per launch, algorithmic bytes per launch)."""
        self.reserve(18)
        out = np.zeros(14)
        self._ck(self.lib.b2d_measure_level1(self._ctx, 0, int(reps), _p(out, _lib.c_f64p)))
        names = ["multi_dot8", "rotate8", "residual", "olsen", "mgs_step", "axpy", "device_copy"]
        return {n: (out[2 * k], out[2 * k + 1]) for k, n in enumerate(names)}

    def fill_op_random(self, side, op_id, seed, amplitude=1.0, symmetric=False):
        self._ck(self.lib.b2d_fill_op_random(self._ctx, side, op_id, int(seed), float(amplitude), int(symmetric)))

    def download_op(self, side, op_id):
        n = int(self.lib.b2d_op_size(self._ctx, side, op_id))
        out = np.empty(max(n, 1))
        self._ck(self.lib.b2d_download_op(self._ctx, side, op_id, _p(out, _lib.c_f64p)))
        return out[:n]


class ProductBlock:
    """Construction of the operators of an ENLARGED block (left child x right child) on the device: the TensorProduct / TensorTrace
    scatter of operatorfunctions.C:19-254 (first device step).  The caller decides which products enter an operator."""

    def __init__(self, left: BlockSpec, right: BlockSpec, q, dims, lmap, rmap, unc_dims, old_to_new, device=0, options=None):
        self.lib = _lib.load()
        self._ctx = C.c_void_p()
        if self.lib.b2d_create(int(device), C.byref(self._ctx)):
            raise B2DError("b2d_create: " + self.lib.b2d_last_error(None).decode())
        for k, v in (options or {}).items():
            self._ck(self.lib.b2d_set_option(self._ctx, k.encode(), float(v)))
        self.op_ids = [[], []]
        for side, blk in enumerate((left, right)):
            bq = np.ascontiguousarray(blk.q, dtype=np.int32).reshape(-1, 3)
            bd = np.ascontiguousarray(blk.dims, dtype=np.int32)
            sites = np.ascontiguousarray(blk.sites, dtype=np.int32)
            self._ck(self.lib.b2d_set_block(self._ctx, side, len(bd), _p(bq, _lib.c_i32p), _p(bd, _lib.c_i32p), int(blk.loop), len(sites), _p(sites, _lib.c_i32p)))
            for op in blk.ops:
                self.op_ids[side].append(SpinBlock._add_op(self, side, op))
        q = np.ascontiguousarray(q, dtype=np.int32).reshape(-1, 3)
        self.dims = np.ascontiguousarray(dims, dtype=np.int32)
        lmap, rmap, unc = (np.ascontiguousarray(x, dtype=np.int32) for x in (lmap, rmap, unc_dims))
        begin = np.zeros(len(old_to_new) + 1, np.int32)
        begin[1:] = np.cumsum([len(x) for x in old_to_new])
        flat = np.ascontiguousarray([u for piece in old_to_new for u in piece], dtype=np.int32)
        self._ck(self.lib.b2d_set_product_stateinfo(self._ctx, len(self.dims), _p(q, _lib.c_i32p), _p(self.dims, _lib.c_i32p), len(unc), _p(lmap, _lib.c_i32p),
                                                     _p(rmap, _lib.c_i32p), _p(unc, _lib.c_i32p), _p(begin, _lib.c_i32p), _p(flat, _lib.c_i32p)))

    def _ck(self, rc):
        if rc:
            raise B2DError("[%d] %s" % (rc, self.lib.b2d_last_error(self._ctx).decode()))

    def set_integrals(self, v1, v2, orbital_irreps, one_tol=1e-15, two_tol=1e-15):
        v1 = np.ascontiguousarray(v1, dtype=np.float64); v2 = np.ascontiguousarray(v2, dtype=np.float64)
        irr = np.ascontiguousarray(orbital_irreps, dtype=np.int32)
        self._ck(self.lib.b2d_set_integrals(self._ctx, len(irr), _p(v1, _lib.c_f64p), _p(v2, _lib.c_f64p), _p(irr, _lib.c_i32p), float(one_tol), float(two_tol)))

    def products(self, optype, orbs, dq, hubbard=False):
        """The planner's product list for one enlarged-block operator: [(left op id or None, left transposed, right op id or None,
        right transposed, scale)] (host work only)."""
        o = np.asarray(list(orbs) + [-1, -1], dtype=np.int32)
        q = np.asarray(dq, dtype=np.int32)
        n = self.lib.b2d_enlarged_op_products(self._ctx, int(optype), len(orbs), _p(o, _lib.c_i32p), _p(q, _lib.c_i32p), int(bool(hubbard)), 0, None, None, None, None)
        if n < 0:
            raise B2DError("[%d] %s" % (-n, self.lib.b2d_last_error(self._ctx).decode()))
        lo, ro, fl, sc = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1))
        self.lib.b2d_enlarged_op_products(self._ctx, int(optype), len(orbs), _p(o, _lib.c_i32p), _p(q, _lib.c_i32p), int(bool(hubbard)), n, _p(lo, _lib.c_i32p),
                                          _p(ro, _lib.c_i32p), _p(fl, _lib.c_i32p), _p(sc, _lib.c_f64p))
        return [(None if lo[k] < 0 else int(lo[k]), bool(fl[k] & 1), None if ro[k] < 0 else int(ro[k]), bool(fl[k] & 2), float(sc[k])) for k in range(n)]

    def build(self, optype, orbs, dq, fermion, hubbard=False, comp=0):
        """Op::build of one operator of the enlarged block, planned on the host and executed on the device: returns its id."""
        o = np.asarray(list(orbs) + [-1, -1], dtype=np.int32)
        q = np.asarray(dq, dtype=np.int32)
        pid = C.c_int(-1)
        self._ck(self.lib.b2d_build_enlarged_op(self._ctx, int(optype), len(orbs), _p(o, _lib.c_i32p), int(comp), _p(q, _lib.c_i32p), int(bool(fermion)),
                                                int(bool(hubbard)), C.byref(pid)))
        return pid.value

    def create(self, dq, fermion):
        """SparseMatrix::allocate on the enlarged block: returns the id of a zero-filled operator with deltaQuantum dq."""
        q = np.asarray(dq, dtype=np.int32)
        pid = C.c_int(-1)
        self._ck(self.lib.b2d_product_op_create(self._ctx, _p(q, _lib.c_i32p), int(bool(fermion)), C.byref(pid)))
        return pid.value

    def accumulate(self, prod_id, left_op, right_op, left_transposed=False, right_transposed=False, scale=1.0):
        """c += scale (a x b); left_op / right_op None = identity on that child (TensorTrace)."""
        self._ck(self.lib.b2d_product_op_accumulate(self._ctx, int(prod_id), -1 if left_op is None else int(left_op), int(bool(left_transposed)),
                                                     -1 if right_op is None else int(right_op), int(bool(right_transposed)), float(scale)))

    def download(self, prod_id):
        nq = len(self.dims)
        n = int(self.lib.b2d_product_op_size(self._ctx, int(prod_id)))
        allowed = np.zeros((nq, nq), np.uint8)
        data = np.empty(max(n, 1))
        self._ck(self.lib.b2d_product_op_download(self._ctx, int(prod_id), _p(allowed, _lib.c_u8p), _p(data, _lib.c_f64p)))
        return allowed.astype(bool), data[:n]

    def kernel_launches(self):
        return int(self.lib.b2d_kernel_launches(self._ctx))

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self.lib.b2d_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def block_spec_from_record(rec, prefix) -> BlockSpec:
    """Build a BlockSpec from a dump record of the reference (oracle/ref_dump.cpp format; tests/golden/*.npz)."""
    q = np.asarray(rec[prefix + "q"], dtype=np.int32).reshape(-1, 3)
    dims = np.asarray(rec[prefix + "dims"], dtype=np.int32)
    blk = BlockSpec(q=q, dims=dims, sites=tuple(int(s) for s in rec[prefix + "sites"]), loop=bool(rec[prefix + "flags"][0]))
    for m in range(int(rec[prefix + "nops"][0])):
        meta = rec["%sop%d.meta" % (prefix, m)]
        norb = int(meta[2])
        blk.ops.append(OperatorSpec(optype=int(meta[0]), orbs=tuple(int(x) for x in meta[3:3 + norb]), comp=int(meta[5]),
                                    dq=(int(meta[6]), int(meta[7]), int(meta[8])), fermion=bool(meta[9]),
                                    allowed=np.asarray(rec["%sop%d.allowed" % (prefix, m)], dtype=np.uint8),
                                    data=np.asarray(rec["%sop%d.data" % (prefix, m)], dtype=np.float64)))
    return blk


def spinblock_from_record(rec, device=0, rank=0, nranks=1, options=None) -> SpinBlock:
    L, R = block_spec_from_record(rec, "L."), block_spec_from_record(rec, "R.")
    norbs = len(rec["spin_orbs_symmetry"]) // 2 if "spin_orbs_symmetry" in rec else None
    return SpinBlock(L, R, tuple(int(x) for x in rec["psi_dq"]), core_energy=float(rec["meta_f"][3]), hubbard=int(rec["meta"][7]) == HUBBARD,
                     norbs=norbs, device=device, rank=rank, nranks=nranks, options=options)


class GuessTransform:
    """Guess wavefunction of the next block iteration: GuessWave::transform_previous_wavefunction (guess_wavefunction.C:524-636,
    two-dot branch) planned on the host (b2d_guess_plan) and executed on the device (b2d_guess_transform).

    `tables` maps the StateInfo names of b2d_guess_desc ("sys", "dot", "left", "right", "oldleft", "oldright", "env") to dicts with
    the reference's tables: "q", "dims", and where the transform reads them "new_quanta_map", "unc.q", "unc.dims", "unc.lmap",
    "unc.rmap", "old_to_new", "old_to_new_begin"."""

    NAMES = ("sys", "dot", "left", "right", "oldleft", "oldright", "env")
    NAMES_BY_MODE = {0: NAMES, 1: ("sys", "dot", "left", "right", "oldleft", "oldright", "oldcol"), 2: ("left", "right", "oldleft", "oldcol"),
                     3: ("left", "right", "oldleft", "oldcol"), 4: ("left", "sys", "dot", "right", "oldleft", "oldcol")}

    def __init__(self, dq, tables, old_allowed, lrot_cols, rrot_cols, device=0, ctx=None, mode=0):
        """mode 0: two-dot step; 1: one-dot, dot on the system side ("oldright" = the reference's newenvstateinfo); 2: one-dot, dot on
        the environment side (see b2d_guess_desc)."""
        self.lib = _lib.load()
        self._own = ctx is None
        if ctx is None:
            self._ctx = C.c_void_p()
            if self.lib.b2d_create(int(device), C.byref(self._ctx)):
                raise B2DError("b2d_create: " + self.lib.b2d_last_error(None).decode())
        else:
            self._ctx = ctx
        self._keep = []     # the arrays the descriptor points into

        def arr(x, dt=np.int32):
            a = np.ascontiguousarray(x, dtype=dt)
            self._keep.append(a)
            return a

        def si(t):
            s = _lib.StateInfoC()
            s.nq = len(t["dims"])
            s.q = _p(arr(t["q"]).reshape(-1), _lib.c_i32p)
            s.dims = _p(arr(t["dims"]), _lib.c_i32p)
            nqm = t.get("new_quanta_map")
            s.new_quanta_map = _p(arr(nqm), _lib.c_i32p) if nqm is not None and len(nqm) else None
            if "unc.dims" in t and "old_to_new_begin" in t:
                s.nunc = len(t["unc.dims"])
                s.unc_q = _p(arr(t["unc.q"]).reshape(-1), _lib.c_i32p)
                s.unc_dims = _p(arr(t["unc.dims"]), _lib.c_i32p)
                s.unc_left = _p(arr(t["unc.lmap"]), _lib.c_i32p)
                s.unc_right = _p(arr(t["unc.rmap"]), _lib.c_i32p)
                s.old_to_new_begin = _p(arr(t["old_to_new_begin"]), _lib.c_i32p)
                s.old_to_new = _p(arr(t["old_to_new"]), _lib.c_i32p)
            else:
                s.nunc = 0
            return s

        d = _lib.GuessDescC()
        for k in range(3):
            d.dq[k] = int(dq[k])
        d.mode = int(mode)
        for name in self.NAMES_BY_MODE[int(mode)]:
            setattr(d, name, si(tables[name]))
        d.old_allowed = _p(arr(old_allowed, np.uint8).reshape(-1), _lib.c_u8p)
        d.lrot_cols = _p(arr(lrot_cols), _lib.c_i32p) if lrot_cols is not None else None
        d.rrot_cols = _p(arr(rrot_cols), _lib.c_i32p) if rrot_cols is not None else None
        self._desc = d
        out = np.zeros(8)
        self._ck(self.lib.b2d_guess_plan(self._ctx, C.byref(d), _p(out, _lib.c_f64p), 8))
        self.old_size, self.lrot_size, self.rrot_size, self.trial_size = (int(x) for x in out[:4])
        self.flops, self.shuffle_bytes, self.shuffle_tasks, self.shuffle_rounds = float(out[4]), float(out[5]), int(out[6]), int(out[7])

    def _ck(self, rc):
        if rc:
            raise B2DError("[%d] %s" % (rc, self.lib.b2d_last_error(self._ctx).decode()))

    def export(self, what):
        """Raw bytes of one descriptor table of the plan (b2d_guess_plan_export)."""
        n = self.lib.b2d_guess_plan_export(self._ctx, int(what), None, 0)
        if n < 0:
            raise B2DError("b2d_guess_plan_export(%d)" % what)
        buf = np.zeros(max(int(n), 1), np.uint8)
        self.lib.b2d_guess_plan_export(self._ctx, int(what), buf.ctypes.data_as(C.c_void_p), int(n))
        return buf[:n]

    def transform(self, old_wave, left_rot=None, right_rot=None, dst_slot=-1, download=True):
        """The trial vector, flat in FlattenInto order (and / or left in wavefunction slot `dst_slot` of a planned context)."""
        ow, lr, rr = (np.ascontiguousarray(x if x is not None else np.zeros(0), dtype=np.float64) for x in (old_wave, left_rot, right_rot))
        assert ow.size == self.old_size and lr.size == self.lrot_size and rr.size == self.rrot_size
        out = np.zeros(max(self.trial_size, 1)) if download else None
        self._ck(self.lib.b2d_guess_transform(self._ctx, _p(ow, _lib.c_f64p), _p(lr, _lib.c_f64p), _p(rr, _lib.c_f64p), int(dst_slot),
                                              _p(out, _lib.c_f64p) if download else None))
        return out[:self.trial_size] if download else None

    def kernel_launches(self):
        return int(self.lib.b2d_kernel_launches(self._ctx))

    def close(self):
        if self._own and self._ctx:
            self.lib.b2d_destroy(self._ctx)
        self._ctx = None
