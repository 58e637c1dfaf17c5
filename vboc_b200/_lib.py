"""ctypes binding of libvboc_b200.so (include/vboc_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `python -m vboc_b200.build`.  There is
no CPU fallback: if the shared object is missing, or no CUDA device is visible when a solver is
created, the calls raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# VBOC_LIB: kernel-tuning experiments load a differently compiled build of the same library (tools/build_variant.sh)
LIB_PATH = os.environ.get("VBOC_LIB") or os.path.join(_HERE, "libvboc_b200.so")

FAMILY_VBOC, FAMILY_AL, FAMILY_MPC = 0, 1, 2
MODE_SQP, MODE_RTI = 0, 1
ERR_ARG, ERR_UNSUPPORTED, ERR_CUDA = -1, -2, -3

EXPORTS = (
    "vboc_default_opts", "vboc_create", "vboc_destroy", "vboc_set_opts", "vboc_set_stream",
    "vboc_solve_batch", "vboc_upload", "vboc_solve_resident", "vboc_solve_resident_async", "vboc_sync",
    "vboc_download", "vboc_last_kernel_ms", "vboc_export_multipliers", "vboc_download_multipliers",
    "vboc_set_mpc", "vboc_set_mpc_reference", "vboc_download_mpc_multipliers", "vboc_set_mpc_rows", "vboc_download_mpc_rows", "vboc_set_cartesian", "vboc_set_mpc_velnorm_start", "vboc_set_guess_network", "vboc_download_guess",
    "vboc_stream_create", "vboc_stream_destroy", "vboc_stream_set_opts", "vboc_stream_free_slots",
    "vboc_stream_pending", "vboc_stream_submit", "vboc_stream_poll", "vboc_stream_fetch", "vboc_stream_sim_step",
    "vboc_datagen_create", "vboc_datagen_destroy", "vboc_datagen_set_opts", "vboc_datagen_run", "vboc_datagen_last_kernel_ms",
    "vboc_testdata_run",
    "vboc_pool_create", "vboc_pool_destroy", "vboc_pool_upload", "vboc_pool_size", "vboc_pool_score", "vboc_pool_select",
    "vboc_pool_remove_selected", "vboc_pool_download", "vboc_pool_download_scores", "vboc_pool_last_score_ms",
    "vboc_sim_step", "vboc_mlp_create", "vboc_mlp_destroy", "vboc_mlp_forward", "vboc_mlp_last_kernel_ms", "vboc_fp64_peak", "vboc_last_error", "vboc_version",
)


class Opts(C.Structure):
    """vboc_opts"""
    _fields_ = [
        ("tol_stat", C.c_double), ("tol_eq", C.c_double), ("tol_ineq", C.c_double), ("tol_comp", C.c_double),
        ("max_iter", C.c_int), ("levenberg_marquardt", C.c_double),
        ("alpha_min", C.c_double), ("alpha_reduction", C.c_double), ("globalization", C.c_int),
        ("qp_tol_stat", C.c_double), ("qp_tol_eq", C.c_double), ("qp_tol_ineq", C.c_double),
        ("qp_tol_comp", C.c_double), ("qp_iter_max", C.c_int),
        ("qp_mu0", C.c_double), ("qp_alpha_min", C.c_double), ("qp_reg_prim", C.c_double),
        ("qp_lam_min", C.c_double), ("qp_t_min", C.c_double), ("qp_tau_min", C.c_double),
    ]


class Stats(C.Structure):
    """vboc_stats"""
    _fields_ = [
        ("status", C.c_int), ("sqp_iter", C.c_int), ("qp_iter", C.c_int), ("ls_evals", C.c_int),
        ("qp_status", C.c_int), ("pad_", C.c_int), ("cost", C.c_double),
        ("res_stat", C.c_double), ("res_eq", C.c_double), ("res_ineq", C.c_double), ("res_comp", C.c_double),
    ]


class DgStats(C.Structure):
    """vboc_dg_stats"""
    _fields_ = [("status", C.c_int), ("n_rows", C.c_int), ("solves", C.c_int), ("converged", C.c_int),
                ("sim_steps", C.c_int), ("sqp_iter", C.c_int), ("qp_iter", C.c_int), ("t_done_us", C.c_int)]


DG_ROWS_MAX = 258


class VbocError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libvboc_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; "
                "g.build()').  vboc_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
        L.vboc_default_opts.argtypes = [C.c_int, C.POINTER(Opts)]
        L.vboc_default_opts.restype = None
        L.vboc_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
        L.vboc_destroy.argtypes = [vp]
        L.vboc_destroy.restype = None
        L.vboc_set_opts.argtypes = [vp, C.POINTER(Opts)]
        L.vboc_set_stream.argtypes = [vp, vp]
        prob = [dp] * 13  # x_guess, u_guess, p, lbx0, ubx0, lbx, ubx, lbxN, ubxN, lbu, ubu, C0 + Tf below
        L.vboc_upload.argtypes = [vp, C.c_int, ip] + [dp] * 12 + [C.c_double]
        L.vboc_solve_batch.argtypes = [vp, C.c_int, C.c_int, ip] + [dp] * 12 + [C.c_double, dp, dp, C.POINTER(Stats)]
        L.vboc_solve_resident.argtypes = [vp, C.c_int]
        L.vboc_solve_resident_async.argtypes = [vp, C.c_int]
        L.vboc_sync.argtypes = [vp]
        L.vboc_download.argtypes = [vp, dp, dp, C.POINTER(Stats)]
        fp_ = C.POINTER(C.c_float)
        L.vboc_set_mpc.argtypes = [vp, C.c_int] + [fp_] * 6 + [C.c_double] * 5 + [dp, dp]
        L.vboc_set_mpc_reference.argtypes = [vp, C.c_int, dp, dp]
        L.vboc_download_mpc_multipliers.argtypes = [vp, dp]
        L.vboc_set_mpc_rows.argtypes = [vp, C.c_int, dp]
        L.vboc_download_mpc_rows.argtypes = [vp, dp]
        L.vboc_set_cartesian.argtypes = [vp, C.c_int] + [C.c_double] * 4
        L.vboc_set_mpc_velnorm_start.argtypes = [vp, C.c_int]
        L.vboc_set_guess_network.argtypes = [vp, C.c_int, C.c_int] + [fp_] * 6 + [C.c_double, C.c_double]
        L.vboc_download_guess.argtypes = [vp, dp]
        L.vboc_export_multipliers.argtypes = [vp, C.c_int]
        L.vboc_download_multipliers.argtypes = [vp, dp, dp]
        L.vboc_last_kernel_ms.argtypes = [vp]
        L.vboc_last_kernel_ms.restype = C.c_double
        L.vboc_sim_step.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, C.c_double, dp]
        L.vboc_stream_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
        L.vboc_stream_destroy.argtypes = [vp]
        L.vboc_stream_destroy.restype = None
        L.vboc_stream_set_opts.argtypes = [vp, C.POINTER(Opts)]
        L.vboc_stream_free_slots.argtypes = [vp]
        L.vboc_stream_pending.argtypes = [vp]
        L.vboc_stream_submit.argtypes = [vp, C.c_int, C.c_int, ip] + [dp] * 12 + [C.c_double, ip]
        L.vboc_stream_poll.argtypes = [vp, C.c_int, ip]
        L.vboc_stream_fetch.argtypes = [vp, C.c_int, dp, dp, C.POINTER(Stats)]
        L.vboc_stream_sim_step.argtypes = [vp, C.c_int, dp, dp, C.c_double, dp]
        L.vboc_datagen_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
        L.vboc_datagen_destroy.argtypes = [vp]
        L.vboc_datagen_destroy.restype = None
        L.vboc_datagen_set_opts.argtypes = [vp, C.POINTER(Opts)]
        L.vboc_datagen_run.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_double, ip, dp, dp, dp, dp, dp, C.c_longlong,
                                       C.POINTER(C.c_longlong), C.POINTER(DgStats)]
        L.vboc_testdata_run.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_int, dp, dp, dp, dp, C.POINTER(DgStats)]
        L.vboc_datagen_last_kernel_ms.argtypes = [vp]
        L.vboc_datagen_last_kernel_ms.restype = C.c_double
        fp = C.POINTER(C.c_float)
        L.vboc_mlp_create.argtypes = [C.c_int] * 5 + [fp] * 6 + [C.POINTER(vp)]
        L.vboc_mlp_destroy.argtypes = [vp]
        L.vboc_mlp_destroy.restype = None
        L.vboc_mlp_forward.argtypes = [vp, C.c_int, fp, C.c_int, C.c_double, C.c_double, C.c_double, fp, fp, ip]
        L.vboc_mlp_last_kernel_ms.argtypes = [vp]
        L.vboc_mlp_last_kernel_ms.restype = C.c_double
        llp = C.POINTER(C.c_longlong)
        L.vboc_pool_create.argtypes = [C.c_int, C.c_int, C.c_longlong, C.POINTER(vp)]
        L.vboc_pool_destroy.argtypes = [vp]
        L.vboc_pool_destroy.restype = None
        L.vboc_pool_upload.argtypes = [vp, C.c_longlong, fp]
        L.vboc_pool_size.argtypes = [vp]
        L.vboc_pool_size.restype = C.c_longlong
        L.vboc_pool_score.argtypes = [vp, vp, C.c_double, C.c_double]
        L.vboc_pool_select.argtypes = [vp, C.c_int, llp, fp, fp]
        L.vboc_pool_remove_selected.argtypes = [vp]
        L.vboc_pool_download.argtypes = [vp, fp]
        L.vboc_pool_download_scores.argtypes = [vp, fp]
        L.vboc_pool_last_score_ms.argtypes = [vp]
        L.vboc_pool_last_score_ms.restype = C.c_double
        L.vboc_fp64_peak.argtypes = [C.c_int, dp]
        L.vboc_last_error.restype = C.c_char_p
        L.vboc_version.restype = C.c_char_p
        del prob
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise VbocError(rc, lib().vboc_last_error().decode())
