"""ctypes binding of libgsgb200.so (declared in include/gsg_b200.h).

The library is the product; this file only loads it and declares argument types.  It raises
loudly if the shared library has not been built -- there is no Python or CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgsgb200.so")


class GsgError(RuntimeError):
    """Raised when a libgsgb200 entry point returns a negative status."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"libgsgb200 error {code}: {msg}")
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  libgsgb200 has no CPU fallback.")
    return C.CDLL(LIB_PATH)


lib = _load()

i32, i64, f64 = C.c_int, C.c_int64, C.c_double
p_i64 = C.POINTER(C.c_int64)
p_f64 = C.POINTER(C.c_double)
vp = C.c_void_p

# name -> (restype, argtypes); every symbol of include/gsg_b200.h
SIGNATURES = {
    "gsg_version": (i32, []),
    "gsg_last_error": (C.c_char_p, []),
    "gsg_device_info": (i32, [i32, C.c_char_p, C.c_size_t]),
    "gsg_launch_count": (i64, []),
    "gsg_get_size": (i32, [i32, i32, i32, i32, p_i64]),
    "gsg_basis_v": (i32, [i32, i32, i32, i32, vp, i64, vp]),
    "gsg_cell_index": (i32, [f64, i32, p_i64]),
    "gsg_basis_tables": (i32, [i32, vp, vp]),
    "gsg_dlf_matrix": (i32, [i32, i32, i32, p_i64, vp, vp, vp]),
    "gsg_block_pattern": (i32, [i32, vp]),
    "gsg_tensor_construct": (i32, [i32, i32, i32, i32, C.POINTER(vp), vp]),
    "gsg_plan_create": (i32, [i32, i32, i32, i32, i64, vp, vp, vp, i32, C.POINTER(vp)]),
    "gsg_plan_destroy": (i32, [vp]),
    "gsg_plan_size": (i32, [vp, p_i64]),
    "gsg_plan_dev_size": (i32, [vp, p_i64]),
    "gsg_tensor_construct_dev": (i32, [vp, C.POINTER(vp), vp]),
    "gsg_pack_dev": (i32, [vp, vp, vp]),
    "gsg_unpack_dev": (i32, [vp, vp, vp]),
    "gsg_plan_set_stream": (i32, [vp, vp]),
    "gsg_plan_sync": (i32, [vp]),
    "gsg_apply_D": (i32, [vp, i32, vp, vp]),
    "gsg_apply_grad": (i32, [vp, vp, vp, vp]),
    "gsg_apply_laplacian": (i32, [vp, vp, vp]),
    "gsg_apply_D_dev": (i32, [vp, i32, f64, vp, f64, vp]),
    "gsg_apply_grad_dev": (i32, [vp, vp, vp, vp]),
    "gsg_apply_laplacian_dev": (i32, [vp, vp, vp, vp]),
    "gsg_apply_dirs_dev": (i32, [vp, vp, C.c_uint, f64, vp, vp]),
    "gsg_rk4_advect": (i32, [vp, vp, vp, f64, i64]),
    "gsg_rk4_advect_dev": (i32, [vp, vp, vp, f64, i64]),
    "gsg_rk4_wave": (i32, [vp, vp, vp, f64, i64]),
    "gsg_rk4_wave_dev": (i32, [vp, vp, vp, f64, i64]),
    "gsg_energy": (i32, [vp, vp, vp, p_f64]),
    "gsg_plan_set_partition": (i32, [vp, i32, i32]),
    "gsg_plan_partition_blocks": (i32, [vp, i32, i32, vp, vp, p_i64, C.POINTER(i32)]),
    "gsg_rk4_taylor_cells_dev": (i32, [vp, vp, i64, vp, vp, vp, vp, vp, f64, f64, f64, f64]),
    "gsg_plan_set_rk4_mode": (i32, [vp, i32]),
    "gsg_plan_set_flat": (i32, [vp, i32]),
    "gsg_plan_flat_active": (i32, [vp, C.POINTER(i32)]),
    "gsg_plan_describe": (i32, [vp, C.c_char_p, C.c_size_t]),
    "gsg_debug_rowtile_program": (i32, [i32, i32, i32, i32, i64, i32, vp, vp, vp, p_i64]),
    "gsg_debug_rowtile_pole_order": (i32, [i32, i32, i32, i32, vp]),
    "gsg_debug_flat_tables": (i32, [i32, i32, i32, i32, i32, vp, vp, p_i64, p_i64]),
    "gsg_ode_create": (i32, [vp, i32, vp, vp, i32, f64, f64, vp, f64, f64, C.POINTER(vp)]),
    "gsg_ode_destroy": (i32, [vp]),
    "gsg_ode_step": (i32, [vp, p_f64, p_f64, C.POINTER(i32)]),
    "gsg_ode_state": (i32, [vp, vp]),
    "gsg_ode_interp": (i32, [vp, f64, vp]),
    "gsg_ode_stats": (i32, [vp, p_i64, p_i64, p_i64]),
    "gsg_vlasov_create": (i32, [vp, vp, vp, vp, vp, C.POINTER(vp), C.POINTER(vp)]),
    "gsg_vlasov_destroy": (i32, [vp]),
    "gsg_vlasov_rhs": (i32, [vp, vp, vp]),
    "gsg_vlasov_v_point": (i32, [vp, i32, vp]),
    "gsg_ode_create_vlasov": (i32, [vp, i32, f64, f64, vp, f64, f64, C.POINTER(vp)]),
    "gsg_mg_create": (i32, [vp, i32, i32, C.POINTER(vp)]),
    "gsg_mg_destroy": (i32, [vp]),
    "gsg_mg_ipc_handle": (i32, [vp, vp]),
    "gsg_mg_connect_ipc": (i32, [vp, vp]),
    "gsg_mg_connect_local": (i32, [C.POINTER(vp), i32]),
    "gsg_mg_set_state": (i32, [vp, vp]),
    "gsg_mg_get_state": (i32, [vp, vp]),
    "gsg_mg_owned_fraction": (i32, [vp, p_f64, p_i64]),
    "gsg_mg_rk4_advect": (i32, [vp, vp, f64, i64]),
    "gsg_mg_rk4_advect_all": (i32, [C.POINTER(vp), i32, vp, f64, i64]),
    "gsg_mg_sync": (i32, [vp]),
    "gsg_rk_stage_dev": (i32, [vp, i64, vp, vp, vp, vp, f64, f64, i32]),
    "gsg_rk_final_dev": (i32, [vp, i64, vp, vp, vp, f64]),
    "gsg_profile_enable": (i32, [vp, i32]),
    "gsg_profile_read": (i32, [vp, p_i64, p_f64, p_f64]),
    "gsg_debug_stamps": (i32, [vp, vp, i32]),
    "gsg_debug_spin": (i32, [vp, i32, i32, i32, i32, i32, i32]),
    "gsg_reconstruct": (i32, [vp, vp, vp, i64, vp]),
    "gsg_reconstruct_dev": (i32, [vp, vp, vp, i64, vp]),
    "gsg_spmv_csc": (i32, [i64, i64, vp, vp, vp, vp, vp]),
    "gsg_csr_create": (i32, [i64, i64, vp, vp, vp, i32, C.POINTER(vp)]),
    "gsg_csr_destroy": (i32, [vp]),
    "gsg_csr_apply": (i32, [vp, vp, vp]),
    "gsg_csr_apply_dev": (i32, [vp, vp, vp, vp]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)      # AttributeError here == missing export: fail loudly
    _fn.restype = _res
    _fn.argtypes = _args


def check(rc: int) -> None:
    if rc != 0:
        raise GsgError(rc, lib.gsg_last_error().decode("utf-8", "replace"))
