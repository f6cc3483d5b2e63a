"""ctypes binding of libemperor_b200.so (include/emperor_b200.h).

The library is built in-tree by `astroemperor_b200/csrc/Makefile`
(`__graft_entry__.build()`).  Loading is lazy so that the host-only modules
(model description, data loading) import on a machine without CUDA, but every
compute entry point goes through `lib()` and raises if the extension is missing:
there is no CPU fallback.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# EMP_B200_LIB: developer override used to A/B kernel variants (scripts/build_variant.sh)
LIB_PATH = os.environ.get("EMP_B200_LIB") or os.path.join(_HERE, "libemperor_b200.so")

_lib = None

EMP_OK = 0
ERROR_NAMES = {-1: "EMP_EINVAL", -2: "EMP_ECUDA", -3: "EMP_ENODEV", -4: "EMP_ENOMEM",
               -5: "EMP_EUNSUPPORTED"}


class EmperorB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {msg}")
        self.code = code


class EmpAmDataC(ctypes.Structure):
    _P = ctypes.c_void_p
    _fields_ = [("n_hipp", ctypes.c_int32), ("n_gost", ctypes.c_int32), ("n_mask2", ctypes.c_int32),
                ("n_mask3", ctypes.c_int32), ("common_t", ctypes.c_double),
                ("time_hipp", _P), ("cpsi_hipp", _P), ("spsi_hipp", _P), ("epoch_hipp", _P),
                ("parf_hipp", _P), ("res_hipp", _P), ("sres_hipp", _P),
                ("time_gost", _P), ("cpsi_gost", _P), ("spsi_gost", _P), ("parf_gost", _P),
                ("idx_mask2", _P), ("idx_mask3", _P), ("gsv2", _P), ("gsv3", _P),
                ("inv_cov", _P), ("log_det_cov", _P), ("astro_gost", _P), ("catalogs", _P)]


_P = ctypes.c_void_p
_I32, _I64 = ctypes.c_int32, ctypes.c_int64
EMP_MAX_PEERS = 16


class EmpPtSweepC(ctypes.Structure):
    """EmpPtSweep of include/emperor_b200.h (layout checked against gcc in tests/test_abi.py)."""
    _fields_ = [("T_loc", _I32), ("W", _I32), ("nsteps", _I32), ("T_all", _I32), ("n_ranks", _I32), ("rank", _I32),
                ("strided", _I32), ("use_graph", _I32),
                ("p", _P), ("logl", _P), ("logp", _P), ("p_alt", _P), ("logl_alt", _P), ("logp_alt", _P),
                ("betas", _P), ("half_idx", _P), ("zz", _P), ("rint", _P), ("factors", _P), ("lnu", _P),
                ("perm", _P), ("lnu_swap", _P), ("accepted", _P), ("n_accepted", _P), ("src", _P), ("n_acc", _P),
                ("adapt", _I32), ("thin", _I32), ("adapt_tau", ctypes.c_double), ("adapt_nu", ctypes.c_double),
                ("sweep_counter", _P), ("step_counter", _P), ("beta_hist", _P), ("nacc_hist", _P), ("smd_hist", _P),
                ("hist_cap", _I64), ("D", _P), ("chain", _P), ("chain_ll", _P), ("chain_lp", _P),
                ("store_cap", _I64), ("store_ring", _I32), ("perm_hot_sorted", _I32),
                ("peer_p", _P * EMP_MAX_PEERS), ("peer_logl", _P * EMP_MAX_PEERS), ("peer_logp", _P * EMP_MAX_PEERS),
                ("logl_all", _P), ("peer_gath", (_P * EMP_MAX_PEERS) * 2)]


# every symbol include/emperor_b200.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("emp_create", ctypes.c_int, [_P, _P, _P, _P, _P, _I64, _P, ctypes.c_int, ctypes.POINTER(_P)]),
    ("emp_destroy", ctypes.c_int, [_P]),
    ("emp_last_error", ctypes.c_char_p, []),
    ("emp_abi_version", ctypes.c_int, []),
    ("emp_stream", ctypes.c_int, [_P, ctypes.POINTER(_P)]),
    ("emp_set_stream", ctypes.c_int, [_P, _P]),
    ("emp_synchronize", ctypes.c_int, [_P]),
    ("emp_attach_sai", ctypes.c_int, [_P, _P, _I32]),
    ("emp_logl_batch", ctypes.c_int, [_P, _P, _I64, _P, _P]),
    ("emp_logl_batch_host", ctypes.c_int, [_P, _P, _I64, _P, _P]),
    ("emp_model_host", ctypes.c_int, [_P, _P, _P, _P]),
    ("emp_kepler_solve_host", ctypes.c_int, [_P, _P, _I64, ctypes.c_int, _P, ctypes.c_int]),
    ("emp_kepler_grid_host", ctypes.c_int, [_P, _P, _I64, ctypes.c_int, _P, _P, _P, ctypes.c_int]),
    ("emp_kepler_grid_table_host", ctypes.c_int, [_P, ctypes.c_double, _I64, _P, _P, _P, ctypes.c_int]),
    ("emp_pt_stretch_step", ctypes.c_int, [_P, _I32, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    ("emp_pt_swap_plan", ctypes.c_int, [_P, _I32, _I32, _P, _P, _P, _P, _P, _P]),
    ("emp_pt_gather_rows", ctypes.c_int, [_P, _I64, _I32, _P, _P, _P, _P, _P, _P, _P]),
    ("emp_nan_count", ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_uint32)]),
    ("emp_draws_create", ctypes.c_int, [_I32, _P, _P, _I32, ctypes.POINTER(_P)]),
    ("emp_draws_destroy", ctypes.c_int, [_P]),
    ("emp_draws_get_state", ctypes.c_int, [_P, _I32, _P, ctypes.POINTER(_I32)]),
    ("emp_draws_sweep", ctypes.c_int, [_P, _P, _I32, _I32, _I32, _P, _P, _P, _P, _P, _I32, _P, _P]),
    ("emp_draws_sweeps", ctypes.c_int, [_P, _I32, _I64, _P, _I32, _I32, _I32, _P, _P, _P, _P, _P, _I32, _P, _P]),
    ("emp_pt_sweep", ctypes.c_int, [_P, ctypes.POINTER(EmpPtSweepC)]),
    ("emp_gather_block_bytes", ctypes.c_int, [_I32, _I32, ctypes.POINTER(_I64)]),
    ("emp_pt_sweep_chunk", ctypes.c_int, [_P, ctypes.POINTER(EmpPtSweepC), ctypes.c_int32, _P, _P, ctypes.c_int64]),
    ("emp_pt_sweep_stretch", ctypes.c_int, [_P, ctypes.POINTER(EmpPtSweepC)]),
    ("emp_pt_sweep_swap", ctypes.c_int, [_P, ctypes.POINTER(EmpPtSweepC)]),
    ("emp_dev_alloc", ctypes.c_int, [ctypes.c_int, _I64, ctypes.POINTER(_P)]),
    ("emp_dev_free", ctypes.c_int, [ctypes.c_int, _P]),
    ("emp_ipc_export", ctypes.c_int, [ctypes.c_int, _P, ctypes.c_char_p]),
    ("emp_ipc_open", ctypes.c_int, [ctypes.c_int, ctypes.c_char_p, ctypes.POINTER(_P)]),
    ("emp_ipc_close", ctypes.c_int, [ctypes.c_int, _P]),
    ("emp_launch_count", ctypes.c_int, [_P, ctypes.POINTER(_I64)]),
    ("emp_graph_captures", ctypes.c_int, [_P, ctypes.POINTER(_I64)]),
    ("emp_set_solver", ctypes.c_int, [_P, ctypes.c_int]),
    ("emp_set_timing", ctypes.c_int, [_P, ctypes.c_int]),
    ("emp_timing_collect", ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_I64)]),
    ("emp_counters", ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_uint64)]),
    ("emp_fp64_peak", ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_double)]),
]


def lib():
    """The loaded C-ABI library; raises (never falls back) if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EmperorB200Error(-3, f"{LIB_PATH} not found: build it with "
                                       "`python -c 'import __graft_entry__ as g; g.build()'` "
                                       "(there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, restype, argtypes in SYMBOLS:
            fn = getattr(L, name)
            fn.restype = restype
            fn.argtypes = argtypes
        from .modelspec import EMP_ABI_VERSION
        if L.emp_abi_version() != EMP_ABI_VERSION:
            raise EmperorB200Error(-1, f"ABI mismatch: library {L.emp_abi_version()} != python {EMP_ABI_VERSION}")
        _lib = L
    return _lib


def check(rc):
    if rc != EMP_OK:
        raise EmperorB200Error(rc, lib().emp_last_error().decode(errors="replace"))
