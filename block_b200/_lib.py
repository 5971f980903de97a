"""ctypes binding of the C ABI in include/block_b200.h (block_b200/lib/libblockb200.so).

The library is built in-tree by block_b200/csrc/build.sh (see __graft_entry__.build).  There is no Python/CPU
implementation behind this module: if the shared object is missing the import fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libblockb200.so")

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_u8p = C.POINTER(C.c_uint8)
c_f64p = C.POINTER(C.c_double)
ctx_p = C.c_void_p



class StateInfoC(C.Structure):
    """b2d_stateinfo (include/block_b200.h)."""
    _fields_ = [("nq", C.c_int32), ("q", c_i32p), ("dims", c_i32p), ("new_quanta_map", c_i32p), ("nunc", C.c_int32), ("unc_q", c_i32p),
                ("unc_dims", c_i32p), ("unc_left", c_i32p), ("unc_right", c_i32p), ("old_to_new_begin", c_i32p), ("old_to_new", c_i32p)]


class GuessDescC(C.Structure):
    """b2d_guess_desc (include/block_b200.h)."""
    _fields_ = [("dq", C.c_int32 * 3), ("mode", C.c_int32), ("sys", StateInfoC), ("dot", StateInfoC), ("left", StateInfoC), ("right", StateInfoC),
                ("oldleft", StateInfoC), ("oldright", StateInfoC), ("env", StateInfoC), ("oldcol", StateInfoC), ("old_allowed", c_u8p),
                ("lrot_cols", c_i32p), ("rrot_cols", c_i32p)]


# name: (restype, [argtypes])  -- one entry per function declared in include/block_b200.h
PROTOTYPES = {
    "b2d_create": (C.c_int, [C.c_int, C.POINTER(ctx_p)]),
    "b2d_destroy": (None, [ctx_p]),
    "b2d_reset": (C.c_int, [ctx_p]),
    "b2d_last_error": (C.c_char_p, [ctx_p]),
    "b2d_abi_version": (C.c_int, []),
    "b2d_set_option": (C.c_int, [ctx_p, C.c_char_p, C.c_double]),
    "b2d_set_block": (C.c_int, [ctx_p, C.c_int, C.c_int, c_i32p, c_i32p, C.c_int, C.c_int, c_i32p]),
    "b2d_add_op": (C.c_int, [ctx_p, C.c_int, C.c_int, C.c_int, c_i32p, C.c_int, c_i32p, C.c_int, c_u8p, c_f64p, C.POINTER(C.c_int)]),
    "b2d_add_op_blocks": (C.c_int, [ctx_p, C.c_int, C.c_int, C.c_int, c_i32p, C.c_int, c_i32p, C.c_int, c_u8p, C.POINTER(c_f64p), C.POINTER(C.c_int)]),
    "b2d_fill_op_random": (C.c_int, [ctx_p, C.c_int, C.c_int, C.c_uint64, C.c_double, C.c_int]),
    "b2d_alloc_ops": (C.c_int, [ctx_p, C.c_int]),
    "b2d_materialise_op": (C.c_int, [ctx_p, C.c_int, C.c_int]),
    "b2d_download_op": (C.c_int, [ctx_p, C.c_int, C.c_int, c_f64p]),
    "b2d_op_size": (C.c_int64, [ctx_p, C.c_int, C.c_int]),
    "b2d_plan": (C.c_int, [ctx_p, c_i32p, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int]),
    "b2d_psi_size": (C.c_int64, [ctx_p]),
    "b2d_psi_padded_size": (C.c_int64, [ctx_p]),
    "b2d_psi_num_blocks": (C.c_int, [ctx_p]),
    "b2d_psi_blocks": (C.c_int, [ctx_p, c_i32p, c_i32p, c_i64p]),
    "b2d_num_terms": (C.c_int, [ctx_p, C.c_int]),
    "b2d_terms": (C.c_int, [ctx_p, C.c_int, c_i32p, c_i32p, c_i32p, c_f64p, c_i32p]),
    "b2d_sigma_flops": (C.c_double, [ctx_p, C.c_int]),
    "b2d_plan_stats": (C.c_int, [ctx_p, c_f64p, C.c_int]),
    "b2d_vec_reserve": (C.c_int, [ctx_p, C.c_int]),
    "b2d_vec_upload": (C.c_int, [ctx_p, C.c_int, c_f64p]),
    "b2d_vec_download": (C.c_int, [ctx_p, C.c_int, c_f64p]),
    "b2d_vec_dot": (C.c_int, [ctx_p, C.c_int, C.c_int, c_f64p]),
    "b2d_vec_axpy": (C.c_int, [ctx_p, C.c_double, C.c_int, C.c_int]),
    "b2d_vec_scale": (C.c_int, [ctx_p, C.c_double, C.c_int]),
    "b2d_vec_copy": (C.c_int, [ctx_p, C.c_int, C.c_int]),
    "b2d_vec_clear": (C.c_int, [ctx_p, C.c_int]),
    "b2d_sigma": (C.c_int, [ctx_p, C.c_int, C.c_int, C.c_int]),
    "b2d_multiplyH_host": (C.c_int, [ctx_p, c_f64p, c_f64p, C.c_int]),
    "b2d_tensor_multiply": (C.c_int, [ctx_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int]),
    "b2d_diagonal": (C.c_int, [ctx_p, C.c_int]),
    "b2d_davidson": (C.c_int, [ctx_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, c_f64p, C.POINTER(C.c_int), c_f64p]),
    "b2d_davidson_lower": (C.c_int, [ctx_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, c_f64p, C.POINTER(C.c_int), c_f64p]),
    "b2d_make_density": (C.c_int, [ctx_p, C.c_int, C.c_int, c_f64p]),
    "b2d_density_size": (C.c_int64, [ctx_p]),
    "b2d_density_download": (C.c_int, [ctx_p, c_f64p]),
    "b2d_density_upload": (C.c_int, [ctx_p, c_f64p]),
    "b2d_diagonalise_dm": (C.c_int, [ctx_p, c_f64p]),
    "b2d_select_states": (C.c_int, [ctx_p, C.c_int, c_i32p, c_f64p]),
    "b2d_rotation_size": (C.c_int64, [ctx_p]),
    "b2d_rotation_download": (C.c_int, [ctx_p, c_f64p]),
    "b2d_rotation_upload": (C.c_int, [ctx_p, c_i32p, c_f64p]),
    "b2d_transform_operators": (C.c_int, [ctx_p]),
    "b2d_rotated_num_sectors": (C.c_int, [ctx_p]),
    "b2d_rotated_sectors": (C.c_int, [ctx_p, c_i32p, c_i32p]),
    "b2d_rotated_op_size": (C.c_int64, [ctx_p, C.c_int]),
    "b2d_rotated_op_download": (C.c_int, [ctx_p, C.c_int, c_u8p, c_f64p]),
    "b2d_rotated_total_size": (C.c_int64, [ctx_p]),
    "b2d_rotated_download_all": (C.c_int, [ctx_p, c_f64p]),
    "b2d_measure_level1": (C.c_int, [ctx_p, C.c_int, C.c_int, c_f64p]),
    "b2d_add_wavefunction_density": (C.c_int, [ctx_p, c_i32p, c_f64p, C.c_double]),
    "b2d_set_product_stateinfo": (C.c_int, [ctx_p, C.c_int, c_i32p, c_i32p, C.c_int, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p]),
    "b2d_product_op_create": (C.c_int, [ctx_p, c_i32p, C.c_int, C.POINTER(C.c_int)]),
    "b2d_product_op_accumulate": (C.c_int, [ctx_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]),
    "b2d_product_stats": (C.c_int, [ctx_p, c_f64p, C.c_int]),
    "b2d_product_op_size": (C.c_int64, [ctx_p, C.c_int]),
    "b2d_product_op_download": (C.c_int, [ctx_p, C.c_int, c_u8p, c_f64p]),
    "b2d_set_integrals": (C.c_int, [ctx_p, C.c_int, c_f64p, c_f64p, c_i32p, C.c_double, C.c_double]),
    "b2d_enlarged_op_products": (C.c_int, [ctx_p, C.c_int, C.c_int, c_i32p, c_i32p, C.c_int, C.c_int, c_i32p, c_i32p, c_i32p, c_f64p]),
    "b2d_build_enlarged_op": (C.c_int, [ctx_p, C.c_int, C.c_int, c_i32p, C.c_int, c_i32p, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "b2d_stash_product": (C.c_int, [ctx_p, C.c_int, C.c_int, C.c_int, c_i32p]),
    "b2d_stash_side": (C.c_int, [ctx_p, C.c_int, C.c_int]),
    "b2d_assemble_big": (C.c_int, [ctx_p]),
    "b2d_renormalise_from": (C.c_int, [ctx_p, C.c_int, C.c_int, c_f64p, C.c_double, C.c_int, C.c_int, C.c_int, C.c_double, c_f64p, c_i32p, c_f64p, C.POINTER(C.c_int)]),
    "b2d_add_onedot_noise": (C.c_int, [ctx_p, C.c_int, C.c_int, C.c_double]),
    "b2d_wavefunction_size": (C.c_int64, [ctx_p, c_i32p]),
    "b2d_tensor_multiply_one_host": (C.c_int, [ctx_p, C.c_int, C.c_int, C.c_int, c_i32p, C.c_double, c_f64p, c_f64p]),
    "b2d_guess_plan": (C.c_int, [ctx_p, C.POINTER(GuessDescC), c_f64p, C.c_int]),
    "b2d_guess_plan_export": (C.c_int64, [ctx_p, C.c_int, C.c_void_p, C.c_int64]),
    "b2d_guess_transform": (C.c_int, [ctx_p, c_f64p, c_f64p, c_f64p, C.c_int, c_f64p]),
    "b2d_cache_put_rotated": (C.c_int, [ctx_p, C.POINTER(C.c_uint64)]),
    "b2d_cache_use": (C.c_int, [ctx_p, C.c_uint64, C.c_int, C.c_int]),
    "b2d_cache_block_info": (C.c_int, [ctx_p, C.c_uint64, c_i32p, c_i32p, c_i32p]),
    "b2d_cache_block_sectors": (C.c_int, [ctx_p, C.c_uint64, c_i32p, c_i32p, c_i32p]),
    "b2d_cache_op_info": (C.c_int, [ctx_p, C.c_uint64, C.c_int, c_i32p, c_i32p, c_i32p, c_i32p, c_i64p]),
    "b2d_cache_download_op": (C.c_int, [ctx_p, C.c_uint64, C.c_int, c_u8p, c_f64p]),
    "b2d_cache_drop": (C.c_int, [ctx_p, C.c_uint64]),
    "b2d_cache_stats": (C.c_int, [ctx_p, c_f64p, C.c_int]),
    "b2d_cache_spill": (C.c_int, [ctx_p, C.c_double]),
    "b2d_nccl_unique_id": (C.c_int, [c_u8p]),
    "b2d_comm_init": (C.c_int, [ctx_p, c_u8p, C.c_int, C.c_int]),
    "b2d_allreduce_slot": (C.c_int, [ctx_p, C.c_int]),
    "b2d_last_timing": (C.c_int, [ctx_p, c_f64p, C.c_int]),
    "b2d_sigma_profile": (C.c_int, [ctx_p, C.c_int, C.c_int, c_f64p]),
    "b2d_kernel_launches": (C.c_int64, [ctx_p]),
    "b2d_sync": (C.c_int, [ctx_p]),
    "b2d_stream": (C.c_void_p, [ctx_p]),
    "b2d_measure_fp64_peak": (C.c_int, [ctx_p, c_f64p, c_f64p]),
}

_lib = None


def load():
    """dlopen the shared library and attach the prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (block_b200/csrc/build.sh). "
                          "There is no CPU fallback for the CUDA path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def nccl_library_path():
    """libnccl.so.2 that ships in the torch wheel (the C library dlopens it lazily; see ctx.cpp)."""
    try:
        import nvidia.nccl  # type: ignore
        for base in list(getattr(nvidia.nccl, "__path__", [])):
            p = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(p):
                return p
    except Exception:
        pass
    return None
