"""ctypes binding of libehic_b200.so (include/ehic_b200.h).  No torch types cross this boundary."""
import ctypes as C
import os

from . import build as _build

OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_NODEVICE, ERR_CAPACITY = range(6)
TERM_LIKELIHOOD, TERM_NONBONDED, TERM_BACKBONE, TERM_SPHERE, TERM_ALL = 1, 2, 4, 8, 15
FLAG_EXACT, FLAG_FP32, FLAG_NB_CELLS = 1, 2, 4
(PARAM_ALPHA, PARAM_NONBONDED_K, PARAM_BACKBONE_K, PARAM_SPHERE_RADIUS, PARAM_SPHERE_K,
 PARAM_SPHERE_CX, PARAM_SPHERE_CY, PARAM_SPHERE_CZ) = range(8)
N_ENERGIES = 4
E_LIKELIHOOD, E_NONBONDED, E_BACKBONE, E_SPHERE = 0, 1, 2, 3     # EHIC_E_* indices of the energy vector

_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)


class ModelDesc(C.Structure):
    _fields_ = [("n_beads", C.c_int32), ("n_structures", C.c_int32), ("n_data", C.c_int64),
                ("data_points", _i64p), ("contact_distances", _dp), ("alpha", C.c_double),
                ("bead_radii", _dp), ("nonbonded_k", C.c_double), ("backbone_k", C.c_double),
                ("bb_lower", _dp), ("bb_upper", _dp), ("mol_ranges", _i32p), ("n_mol_ranges", C.c_int32),
                ("sphere_active", C.c_int32), ("sphere_radius", C.c_double), ("sphere_k", C.c_double),
                ("sphere_center", C.c_double * 3), ("max_replicas", C.c_int32), ("flags", C.c_int32)]


class EhicError(RuntimeError):
    def __init__(self, code, msg):
        super(EhicError, self).__init__("libehic_b200 error %d: %s" % (code, msg))
        self.code = code


#: every symbol include/ehic_b200.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = ["ehic_version", "ehic_last_error", "ehic_device_count", "ehic_model_create", "ehic_model_destroy",
           "ehic_model_set_param", "ehic_model_get_param", "ehic_model_info", "ehic_mock_data", "ehic_evaluate",
           "ehic_hmc_trajectory", "ehic_trajectory_plan", "ehic_select", "ehic_nblist_host", "ehic_mock_data_host", "ehic_evaluate_host",
           "ehic_layout_debug", "ehic_launch_count", "ehic_profile", "ehic_profile_read"]

_lib = None


def lib():
    """Load (building first if stale) the shared library.  Raises if it cannot be built/loaded:
    the product path has no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("EHIC_B200_LIB") or _build.build()      # override: a development build of the same library
    L = C.CDLL(path)
    vp = C.c_void_p
    L.ehic_version.restype = C.c_int
    L.ehic_last_error.restype = C.c_char_p
    L.ehic_device_count.restype = C.c_int
    L.ehic_launch_count.restype = C.c_int64
    L.ehic_model_create.argtypes = [C.POINTER(ModelDesc), C.c_int, C.POINTER(vp)]
    L.ehic_model_destroy.argtypes = [vp]
    L.ehic_model_set_param.argtypes = [vp, C.c_int, C.c_double]
    L.ehic_model_get_param.argtypes = [vp, C.c_int, _dp]
    L.ehic_model_info.argtypes = [vp, _i64p]
    L.ehic_mock_data.argtypes = [vp, C.c_int, vp, vp, vp, vp]
    L.ehic_evaluate.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp]
    L.ehic_hmc_trajectory.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp]
    L.ehic_trajectory_plan.argtypes = [vp, C.c_int, C.c_int, _i64p]
    L.ehic_select.argtypes = [C.c_int, C.c_int64, vp, vp, vp, vp]
    L.ehic_nblist_host.argtypes = [C.c_int, vp, C.c_double, vp, C.c_int64, _i64p, C.c_int]
    L.ehic_mock_data_host.argtypes = [vp, C.c_int, vp, vp, vp]
    L.ehic_evaluate_host.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp]
    L.ehic_layout_debug.argtypes = [C.c_int32, C.c_int64, vp, C.c_int, vp, _i64p, vp]
    L.ehic_profile.argtypes = [vp, C.c_int]
    L.ehic_profile_read.argtypes = [vp, _dp, _i64p]
    _lib = L
    return L


def check(rc):
    if rc != OK:
        raise EhicError(rc, lib().ehic_last_error().decode("utf-8", "replace"))


def loaded_path():
    return _build.LIB if _lib is not None else None
