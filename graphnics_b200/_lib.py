"""ctypes binding of libgraphnics_b200.so (the C ABI declared in include/graphnics_b200.h).

There is NO CPU fallback: if the shared library is missing, or a compute entry point is
called without a CUDA device, this module raises.  The library is built in-tree by
`make -C graphnics_b200/csrc` (or `__graft_entry__.build()`).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgraphnics_b200.so")

c_i32p = C.POINTER(C.c_int32)
c_f64p = C.POINTER(C.c_double)


class GnxError(RuntimeError):
    pass


class gnx_network_t(C.Structure):
    _fields_ = [
        ("V", C.c_int32), ("E", C.c_int32), ("gdim", C.c_int32), ("n", C.c_int32),
        ("B", C.c_int32), ("n_bnd", C.c_int32),
        ("I", C.c_int64),
        ("pos", C.c_void_p), ("edges", C.c_void_p), ("h", C.c_void_p), ("bix", C.c_void_p),
        ("bif_nodes", C.c_void_p), ("inc_ptr", C.c_void_p), ("inc_edge", C.c_void_p),
        ("lam_ptr", C.c_void_p), ("ebif_prefix", C.c_void_p), ("loc", C.c_void_p),
        ("tloc", C.c_void_p), ("rowtab", C.c_void_p), ("erec", C.c_void_p),
    ]


class gnx_mixed_coeffs_t(C.Structure):
    _fields_ = [("a_edge", C.c_void_p), ("k_edge", C.c_void_p), ("sigma", C.c_double), ("a_cell", C.c_void_p)]


MAX_LEVELS = 64
TOP_MAX = 2048
BLK_MAX = 1024                # GNX_BLK_MAX: nodes of a rake block (one CTA, one node per thread)
SUB_PER_AGG = int(os.environ.get("GNX_SCHUR_SUB", "8"))   # fine aggregates per coarse one (1: coarse level only)
AGG_MAX = 128                 # GNX_AGG_MAX: aggregates (= CTAs) of the two-level PCG
TWO_LEVEL_MIN_CORE = int(os.environ.get("GNX_TWO_LEVEL_MIN_CORE", "2048"))   # smaller cores: one CTA or one cluster, plain Jacobi


class gnx_schur_plan_t(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("nl", C.c_int32), ("l_split", C.c_int32), ("nc", C.c_int32),
        ("nnzc", C.c_int32), ("nt", C.c_int32),
        ("lvl_ptr", C.c_int32 * (MAX_LEVELS + 1)),
        ("lvl_nodes", C.c_void_p), ("parent", C.c_void_p), ("pedge", C.c_void_p), ("ch_ptr", C.c_void_p),
        ("ch_idx", C.c_void_p), ("core_nodes", C.c_void_p), ("crp", C.c_void_p), ("ccol", C.c_void_p),
        ("cedge", C.c_void_p), ("top_nodes", C.c_void_p), ("top_of", C.c_void_p), ("level_of", C.c_void_p),
        ("ell_w", C.c_int32), ("ell_R", C.c_int32), ("top_S", C.c_int32), ("na", C.c_int32),
        ("ell_col", C.c_void_p), ("ell_edge", C.c_void_p),
        ("agg_ptr", C.c_void_p), ("col_agg", C.c_void_p), ("cut_ptr", C.c_void_p), ("cut_k", C.c_void_p),
        ("nb", C.c_int32), ("ns", C.c_int32),
        ("blk_ptr", C.c_void_p), ("blk_nodes", C.c_void_p), ("blk_pos", C.c_void_p), ("blk_lv", C.c_void_p),
        ("sub_ptr", C.c_void_p), ("row_sub", C.c_void_p),
        ("l_sync", C.c_int32), ("nxl", C.c_int32), ("ncls", C.c_void_p), ("xlist", C.c_void_p),
        ("top_rec", C.c_void_p),
    ]


class gnx_exchange_t(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("chunk", C.c_int32), ("mode", C.c_int32),
                ("peers_host", C.POINTER(C.c_void_p)), ("epoch", C.c_int64), ("egl", C.c_void_p),
                ("E_glob", C.c_int32), ("B_glob", C.c_int32), ("phase", C.c_int32), ("reserved", C.c_int32)]


_P = C.c_void_p
_I = C.c_int32
_L = C.c_int64
_D = C.c_double
_NET = C.POINTER(gnx_network_t)
_CO = C.POINTER(gnx_mixed_coeffs_t)
_PLAN = C.POINTER(gnx_schur_plan_t)
_EX = C.POINTER(gnx_exchange_t)


class gnx_primal_coeffs_t(C.Structure):
    _fields_ = [("a_const", C.c_double), ("a_cell", C.c_void_p), ("sigma", C.c_double)]


_PCO = C.POINTER(gnx_primal_coeffs_t)

# name -> (restype, argtypes); every symbol of include/graphnics_b200.h
PROTOTYPES = {
    "gnx_last_error": (C.c_char_p, []),
    "gnx_version": (_I, []),
    "gnx_device_ok": (_I, []),
    "gnx_refine_1d": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "gnx_edge_tangents": (_I, [_P, _P, _I, _I, _P, _P, _P]),
    "gnx_tables_work_ints": (_L, [_I, _I]),
    "gnx_vertex_tables": (_I, [_P, _I, _I] + [_P] * 14),
    "gnx_mixed_ndofs": (_L, [_NET]),
    "gnx_mixed_nnz": (_L, [_NET]),
    "gnx_build_pattern_mixed": (_I, [_NET, _P, _P, _P]),
    "gnx_assemble_mixed": (_I, [_NET, _CO, _P, _P, _P]),
    "gnx_assemble_mass_q_mixed": (_I, [_NET, _D, _P, _P, _P]),
    "gnx_assemble_rhs_mixed": (_I, [_NET, _P, _P, _D, _D, _P, _P, _P, _P]),
    "gnx_eval_expr": (_I, [_P, _P, _I, _L, _I, _P, _P, _P, _P]),
    "gnx_project_p1_edges": (_I, [_NET, _P, _P, _P, _P, _P]),
    "gnx_project_end_values": (_I, [_NET, _P, _P, _I, _I, _P, _P]),
    "gnx_dof_coordinates_mixed": (_I, [_NET, _P, _P, _P, _P, _P]),
    "gnx_edge_eliminate_mixed": (_I, [_NET, _CO, _P, _P, _P, _P, _P]),
    "gnx_edge_eliminate_mixed_p2p": (_I, [_NET, _CO, _P, _EX, _P, _P, _P]),
    "gnx_rhs_eliminate_mixed": (_I, [_NET, _CO, _D, _D, _P, _P, _P, _P, _P, _P, _EX, _P]),
    "gnx_step_mixed": (_I, [_NET, _CO, _P, _P, _D, _D, _P, _P, _P, _P, _NET, _PLAN, _P, _P, _P, _P, _P, _D, _I, _I,
                            _EX, c_f64p, c_i32p, _P, _P, _P, _P, _L, _P]),
    "gnx_step_is_resident": (_I, []),
    "gnx_timestep_mixed": (_I, [_NET, _CO, _PLAN, _P, _P, _D, _D, _P, _P, _D, _D, _P, _P, _P, _P, _P, _D, _P, _P, _P, _P,
                                _P, _P, _P, _P, _P, _P, _D, _I, _I, c_f64p, c_i32p, _P]),
    "gnx_timestep_primal": (_I, [_NET, _PCO, _PLAN, _P, _P, _I, _D, _P, _I, _D, _P, _D, _D, _P, _P, _P, _P, _P, _PCO, _P, _P,
                                 _P, _P, _P, _P, _P, _P, _P, _P, _D, _I, _I, c_f64p, c_i32p, _P]),
    "gnx_schur_build": (_I, [_NET, _D, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gnx_schur_work_doubles": (_L, [_PLAN]),
    "gnx_schur_unconverged": (_I, [_PLAN, _P, _P]),
    "gnx_schur_solve": (_I, [_NET, _PLAN, _D, _P, _P, _P, _P, _P, _P, _P, _P, _D, _I, _I, _I, _I, _EX, c_f64p, c_i32p, _P]),
    "gnx_primal_ndofs": (_L, [_NET]),
    "gnx_primal_nnz": (_L, [_NET]),
    "gnx_build_pattern_primal": (_I, [_NET, _P, _P, _P]),
    "gnx_assemble_primal": (_I, [_NET, _PCO, _P, _P]),
    "gnx_assemble_rhs_primal": (_I, [_NET, _P, _P, _I, _D, _P, _I, _D, _P, _P]),
    "gnx_mass_apply_q_primal": (_I, [_NET, _PCO, _P, _P, _P]),
    "gnx_apply_bc_primal": (_I, [_NET, _D, _P, _P, _P, _P]),
    "gnx_edge_eliminate_primal": (_I, [_NET, _PCO, _P, _P, _P, _P, _P]),
    "gnx_edge_backsub_primal": (_I, [_NET, _PCO, _P, _P, _P, _P, _P, _P]),
    "gnx_edge_backsub_mixed": (_I, [_NET, _CO, _P, _P, _P, _P, _P, _P]),
    "gnx_edge_backsub_mixed_const": (_I, [_NET, _CO, _D, _D, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gnx_edge_flux_mixed": (_I, [_NET, _P, _P, _P, _P, _P, _P]),
    "gnx_edge_backsub_mixed_flux": (_I, [_NET, _CO, _D, _D, _P, _P, _P, _P, _P, _P]),
    "gnx_spmv_csr_f64": (_I, [_I, _P, _P, _P, _P, _P, _P]),
    "gnx_reduce_work_doubles": (_L, [_L]),
    "gnx_residual_csr_f64": (_I, [_I, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gnx_axpby_dot": (_I, [_L, _D, _P, _D, _P, _P, _P, _P]),
    "gnx_lincomb": (_I, [_L, _I, c_f64p, C.POINTER(C.c_void_p), _P, _P]),
    "gnx_mass_apply_q_mixed": (_I, [_NET, _D, _P, _P, _P]),
}

_lib = None


def load():
    """dlopen the C-ABI library (no CUDA device needed just to load and list symbols)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GnxError(
            f"{LIB_PATH} not found: build it with `make -C graphnics_b200/csrc` "
            "(graphnics_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().gnx_last_error()
        raise GnxError(f"graphnics_b200 C-ABI call failed (rc={rc}): {msg.decode() if msg else ''}")


def require_device():
    lib = load()
    if not lib.gnx_device_ok():
        raise GnxError("graphnics_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return lib


def ptr(t):
    """device pointer of a torch tensor (or None -> NULL)"""
    return None if t is None else C.c_void_p(t.data_ptr())
