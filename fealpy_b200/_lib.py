"""ctypes binding of libfealpy_b200.so (the C ABI declared in include/fealpy_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is present
when a kernel is requested, the call raises.  PyTorch tensors are the only device
containers; every entry point receives raw `data_ptr()`s and the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FB2_LIB_PATH") or os.path.join(_HERE, "libfealpy_b200.so")   # override: tuning builds (tools/build_variants.sh)

FB2_ERRORS = {1: ValueError, 2: RuntimeError, 3: NotImplementedError, 4: RuntimeError}

_p = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int
_f64 = C.c_double
_sz = C.c_size_t

# name -> (restype, argtypes); must list every symbol of include/fealpy_b200.h
SIGNATURES = {
    "fb2_last_error": (C.c_char_p, []),
    "fb2_version": (_i32, []),
    "fb2_tri_from_box": (_i32, [_p, _i32, _i32, _p, _p, _p]),
    "fb2_tet_from_box": (_i32, [_p, _i32, _i32, _i32, _p, _p, _p]),
    "fb2_entity_workspace_bytes": (_sz, [_i64, _i32]),
    "fb2_build_entities": (_i32, [_p, _i64, _i32, _i32, _i64, _p, _p, _p, _p]),
    "fb2_entities_emit": (_i32, [_p, _i64, _i32, _i32, _p, _p, _p, _p]),
    "fb2_cell_to_dof": (_i32, [_p, _p, _p, _p, _i64, _i32, _i32, _i64, _i64, _i64, _p, _i32, _p, _p]),
    "fb2_tensor_cell_to_dof": (_i32, [_p, _i64, _i32, _i32, _i64, _i32, _p, _p]),
    "fb2_elem_scalar_const": (_i32, [_i32, _i32, _i64, _p, _p, _p, _p, _f64, _p, _f64, _p, _p, _p]),
    "fb2_elem_scalar_quad": (_i32, [_i32, _i32, _i64, _p, _p, _i32, _i32, _p, _p, _i32, _p, _p, _p]),
    "fb2_elem_scalar_const_acc": (_i32, [_i32, _i32, _i64, _p, _p, _p, _p, _f64, _p, _f64, _p, _p, _i32, _p]),
    "fb2_elem_scalar_quad_fused": (_i32, [_i32, _i32, _i64, _p, _p, _i32, _i32, _p, _p, _i32, _p, _p, _p, _f64, _p, _f64, _p, _p, _i32, _p]),
    "fb2_elem_elasticity": (_i32, [_i32, _i32, _i64, _p, _p, _p, _f64, _f64, _f64, _i32, _p, _p]),
    "fb2_cell_gradients": (_i32, [_i32, _i64, _p, _p, _p, _p]),
    "fb2_grad_basis": (_i32, [_i32, _i64, _i32, _i32, _p, _p, _p, _p]),
    "fb2_coo_keys_from_c2d": (_i32, [_p, _p, _i64, _i32, _i32, _i32, _p, _p]),
    "fb2_coo_keys_from_coo": (_i32, [_p, _p, _i32, _i64, _i32, _p, _p]),
    "fb2_coo_workspace_bytes": (_sz, [_i64]),
    "fb2_coo_symbolic": (_i32, [_p, _p, _i64, _i32, _p, _p, _p]),
    "fb2_coo_fill": (_i32, [_p, _i64, _i32, _i64, _p, _p, _p, _i32, _p, _p]),
    "fb2_coo_reduce": (_i32, [_p, _p, _i64, _p, _p, _p]),
    "fb2_sym_workspace_bytes": (_sz, [_i64, _i32, _i64]),
    "fb2_bc_to_points": (_i32, [_i32, _i64, _i32, _p, _p, _p, _p, _p]),
    "fb2_assemble_elasticity_p1": (_i32, [_i32, _i64, _p, _p, _i32, _i64, _f64, _f64, _f64, _f64, _p, _p, _p, _i32, _p, _i32, _p, _p, _i32,
                                          _i32, _p, _p, _p]),
    "fb2_sym_count": (_i32, [_p, _i64, _i32, _i64, _p, _p, _p, _p, _p, _p, _p, _p]),
    "fb2_sym_fill": (_i32, [_p, _i64, _i32, _i64, _p, _p, _p, _p, _p, _i32, _p, _p]),
    "fb2_slot_stride": (_i32, [_i32, _i32]),
    "fb2_assemble_scalar_const": (_i32, [_i32, _i32, _i64, _i64, _p, _p, _p, _p, _p, _i32, _p, _i32, _p, _i32, _i32, _p, _p,
                                         _f64, _p, _f64, _p, _p, _p]),
    "fb2_assemble_from_ke": (_i32, [_i64, _i32, _i32, _i32, _i64, _p, _p, _p, _p, _i32, _p, _i32, _p, _p, _i32, _i32, _p, _p]),
    "fb2_asm4_workspace_bytes": (_sz, [_i32]),
    "fb2_asm4_plan_count": (_i32, [_i32, _p, _p, _p, _p, _i32, _p, _p, _p, _p]),
    "fb2_asm4_plan_fill": (_i32, [_i32, _p, _p, _p, _p, _i32, _p, _p, _p, _p, _i32, _p]),
    "fb2_assemble_scalar_const_v4": (_i32, [_i32, _i32, _i64, _p, _p, _p, _p, _i32, _i32, _i32, _p, _p, _p, _i32, _p, _p,
                                            _f64, _p, _f64, _p, _p, _p, _p]),
    "fb2_expand_pattern": (_i32, [_i64, _i32, _i32, _p, _p, _p, _p, _p]),
    "fb2_partial_workspace_bytes": (_sz, []),
    "fb2_spmv_plan_blocks": (_i32, [_i64, _i32]),
    "fb2_spmv_plan_build": (_i32, [_i64, _p, _i32, _p, _i64, _p, _p]),
    "fb2_csr_spmv": (_i32, [_i64, _i64, _p, _p, _p, _p, _p, _p, _i32, _i32, _p]),
    "fb2_csr_spmm": (_i32, [_i64, _p, _p, _p, _p, _p, _i32, _p]),
    "fb2_dot": (_i32, [_i64, _p, _p, _p, _p, _p]),
    "fb2_cg_workspace_bytes": (_sz, [_i64, _i64]),
    "fb2_cg": (_i32, [_i64, _i64, _p, _p, _p, _p, _p, _p, _f64, _f64, _i32, _i32, _p, _i32, _i32, _p, _p, _p, _p]),
    "fb2_cg_init": (_i32, [_p, _f64, _f64, _i32, _f64, _f64, _p]),
    "fb2_cg_residual": (_i32, [_i64, _i64, _p, _p, _p, _p, _p, _p, _p, _i32, _i32, _p]),
    "fb2_cg_start": (_i32, [_i64, _p, _p, _p, _p, _p, _p, _p]),
    "fb2_cg_spmv_dot": (_i32, [_i64, _i64, _p, _p, _p, _p, _p, _p, _i32, _i32, _p, _p, _p, _p]),
    "fb2_cg_update_xr": (_i32, [_i64, _p, _p, _p, _p, _p, _p, _p, _i32, _p, _p]),
    "fb2_box_edges_before": (_i64, [_i32, _i32, _i32, _i32, _i32, _i32]),
    "fb2_tet_box_slab": (_i32, [_p, _i32, _i32, _i32, _i32, _i32, _i32, _p, _p, _p, _p]),
    "fb2_cg_finalize": (_i32, [_p, _p]),
    "fb2_cg_update_p": (_i32, [_i64, _p, _p, _p, _p, _p]),
    "fb2_gather_f64": (_i32, [_i64, _p, _p, _p, _p]),
    "fb2_peer_ctrl_bytes": (_i32, []),
    "fb2_cg_spmv_dot_ranges": (_i32, [_i64, _i64, _p, _p, _p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _p, _i32, _p, _p, _p]),
    "fb2_peer_allreduce": (_i32, [_p, _p, _i32, _i32, _i32, _p, _p, _p, _p, _i32, _p, _p]),
    "fb2_peer_wait_halo": (_i32, [_p, _i32, _p, _p, _p, _p]),
    "fb2_cg_update_p_push": (_i32, [_p, _p, _p, _p, _p, _i32, _p, _p, _p, _p, _i32, _p, _i32, _p, _p, _p]),
    "fb2_bcg_dots": (_i32, [_i64, _i32, _p, _p, _p, _p, _p]),
    "fb2_bcg_update_xr": (_i32, [_i64, _i32, _p, _p, _p, _p, _p, _p, _p, _p]),
    "fb2_bcg_update_p": (_i32, [_i64, _i32, _p, _p, _p, _p, _p, _p, _p]),
    "fb2_bcg_check": (_i32, [_i32, _p, _f64, _f64, _i32, _p, _p]),
    "fb2_elem_source": (_i32, [_i32, _i64, _i32, _i32, _p, _p, _p, _i32, _f64, _p, _p, _p]),
    "fb2_gather_vector": (_i32, [_i64, _p, _p, _p, _p, _p]),
    "fb2_matfree_apply": (_i32, [_i64, _i32, _p, _p, _p, _p, _p, _p, _p]),
    "fb2_adjacency_workspace_bytes": (_sz, [_i64]),
    "fb2_adjacency": (_i32, [_p, _i64, _i32, _i64, _p, _p, _p, _p]),
    "fb2_matfree_scalar_const": (_i32, [_i32, _i32, _i64, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _f64, _p, _f64, _p, _p, _p, _p, _p]),
    "fb2_pair_positions": (_i32, [_i64, _p, _p, _p]),
    "fb2_bc_workspace_bytes": (_sz, [_i64]),
    "fb2_bc_matrix_count": (_i32, [_i64, _p, _p, _p, _p, _p, _p, _p]),
    "fb2_bc_matrix_fill": (_i32, [_i64, _p, _p, _p, _p, _p, _p, _p, _p]),
    "fb2_bc_vector": (_i32, [_i64, _p, _p, _p, _p]),
    "fb2_sort_workspace_bytes": (_sz, [_i64]),
    "fb2_sort_pairs": (_i32, [_p, _p, _i32, _i64, _i32, _p, _p]),
    "fb2_scan_workspace_bytes": (_sz, [_i64]),
    "fb2_exclusive_scan_i32": (_i32, [_p, _p, _i64, _p, _p]),
}

_lib = None


def load():
    """Load the shared library (no GPU needed for loading / symbol checks)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"fealpy_b200: {LIB_PATH} not found -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C fealpy_b200/csrc`).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so misses a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("fealpy_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def ptr(t):
    """device (or host) pointer of a tensor, None -> NULL"""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def check(status: int):
    if status != 0:
        msg = load().fb2_last_error().decode(errors="replace")
        raise FB2_ERRORS.get(status, RuntimeError)(f"fealpy_b200: {msg}")


def call(name, *args):
    """invoke a status-returning entry point and raise on failure"""
    lib = load()
    check(getattr(lib, name)(*args))


def workspace(nbytes: int, device):
    return torch.empty(int(nbytes), dtype=torch.uint8, device=device)


_partial_ws = {}


def partial_ws(device):
    """zero-initialised scratch for the deterministic reductions (one per device + stream)"""
    key = (torch.device(device).index, torch.cuda.current_stream().cuda_stream)
    if key not in _partial_ws:
        _partial_ws[key] = torch.zeros(int(load().fb2_partial_workspace_bytes()), dtype=torch.uint8, device=device)
    return _partial_ws[key]
