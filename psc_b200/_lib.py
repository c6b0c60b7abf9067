"""ctypes loader for libpsc_b200.so (the C ABI declared in include/psc_b200.h).

There is no fallback of any kind: if the shared library is missing or a call
fails, an exception is raised."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
MAX_KINDS = 10

i3 = C.c_int * 3
d3 = C.c_double * 3


class GridDesc(C.Structure):
    """psc_b200_grid_desc"""
    _fields_ = [
        ("gdims", i3), ("np", i3),
        ("length", d3), ("corner", d3),
        ("dt", C.c_double), ("fnqs", C.c_double), ("eta", C.c_double),
        ("n_kinds", C.c_int),
        ("q", C.c_double * MAX_KINDS), ("m", C.c_double * MAX_KINDS),
        ("bc_fld_lo", i3), ("bc_fld_hi", i3), ("bc_prt_lo", i3), ("bc_prt_hi", i3),
        ("deposit", C.c_int),
        ("rank", C.c_int), ("n_ranks", C.c_int),
        ("n_patches_by_rank", C.POINTER(C.c_int)),
        ("device", C.c_int),
        ("max_n_prts", C.c_uint64),
    ]


class StepParams(C.Structure):
    """psc_b200_step_params"""
    _fields_ = [
        ("sort", C.c_int), ("marder_loop", C.c_int), ("marder_diffusion", C.c_double),
        ("push_fields", C.c_int), ("checks", C.c_int), ("energies", C.c_int),
    ]


class CollisionParams(C.Structure):
    """psc_b200_collision_params"""
    _fields_ = [
        ("interval", C.c_int), ("nu", C.c_double), ("cori", C.c_double), ("rng", C.c_int),
        ("seed", C.c_uint64), ("step", C.c_uint64),
    ]


class HeatingParams(C.Structure):
    """psc_b200_heating_params"""
    _fields_ = [
        ("zl", C.c_double), ("zh", C.c_double), ("xc", C.c_double), ("yc", C.c_double), ("rH", C.c_double),
        ("T", C.c_double * MAX_KINDS), ("Mi", C.c_double), ("n_kinds", C.c_int), ("interval", C.c_int),
        ("seed", C.c_uint64), ("step", C.c_uint64),
    ]


class JPath(C.Structure):
    """psc_b200_jpath"""
    _fields_ = [
        ("patch", C.c_int), ("lg", C.c_int * 3), ("xm", C.c_float * 3), ("xp", C.c_float * 3),
        ("v", C.c_float * 3), ("qni_wni", C.c_float),
    ]


class PscB200Error(RuntimeError):
    pass


_lib = None


def lib_path():
    name = os.environ.get("PSC_B200_LIB", "libpsc_b200.so")
    if os.path.isabs(name):
        return name
    return os.path.join(HERE, "lib", name)


def load():
    """loads the CUDA library; raises if it has not been built"""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise PscB200Error(
            f"{path} not found: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()' or make -C psc_b200/csrc)")
    L = C.CDLL(path)
    P = C.c_void_p
    CTX = C.c_void_p
    L.psc_b200_last_error.restype = C.c_char_p
    L.psc_b200_version.restype = C.c_char_p
    sigs = {
        "psc_b200_create": [C.POINTER(GridDesc), C.POINTER(CTX)],
        "psc_b200_sync": [CTX],
        "psc_b200_n_patches": [CTX],
        "psc_b200_patch_begin": [CTX],
        "psc_b200_get_ldims": [CTX, i3, i3],
        "psc_b200_mprts_set": [CTX, P, P],
        "psc_b200_mprts_inject": [CTX, P, P],
        "psc_b200_mprts_size": [CTX, C.POINTER(C.c_uint64)],
        "psc_b200_mprts_size_by_patch": [CTX, P],
        "psc_b200_mprts_get": [CTX, P, P],
        "psc_b200_mprts_setup_thermal": [CTX, C.c_int, P, C.c_uint64],
        "psc_b200_mprts_setup_thermal_by_patch": [CTX, P, P, C.c_uint64],
        "psc_b200_mflds_create": [CTX, C.c_int, C.POINTER(C.c_int)],
        "psc_b200_mflds_upload": [CTX, C.c_int, C.c_int, C.c_int, P],
        "psc_b200_mflds_download": [CTX, C.c_int, C.c_int, C.c_int, P],
        "psc_b200_mflds_zero": [CTX, C.c_int, C.c_int, C.c_int],
        "psc_b200_mflds_fill": [CTX, C.c_int, C.c_int, C.c_float],
        "psc_b200_push_mprts": [CTX],
        "psc_b200_sort": [CTX],
        "psc_b200_bnd_particles": [CTX],
        "psc_b200_bnd_add_ghosts": [CTX, C.c_int, C.c_int, C.c_int],
        "psc_b200_bnd_fill_ghosts": [CTX, C.c_int, C.c_int, C.c_int],
        "psc_b200_bndf_fill_ghosts_E": [CTX],
        "psc_b200_bndf_fill_ghosts_H": [CTX],
        "psc_b200_bndf_add_ghosts_J": [CTX],
        "psc_b200_push_E": [CTX, C.c_double],
        "psc_b200_push_H": [CTX, C.c_double],
        "psc_b200_marder": [CTX, C.c_double, C.c_int],
        "psc_b200_moment_rho_1st_nc": [CTX, C.c_int],
        "psc_b200_moment_n_comps": [CTX, C.c_int],
        "psc_b200_selftest_math": [CTX, C.POINTER(C.c_uint64)],
        "psc_b200_moment_1st": [CTX, C.c_int, C.c_int],
        "psc_b200_check_continuity_begin": [CTX],
        "psc_b200_check_continuity_end": [CTX, C.POINTER(C.c_double)],
        "psc_b200_check_gauss": [CTX, C.POINTER(C.c_double)],
        "psc_b200_collide": [CTX, C.POINTER(CollisionParams), P],
        "psc_b200_heating_spot_foil": [CTX, C.POINTER(HeatingParams), P],
        "psc_b200_deposit_j": [CTX, P, C.c_uint64],
        "psc_b200_mflds_add": [CTX, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int],
        "psc_b200_mflds_scale": [CTX, C.c_int, C.c_int, C.c_int, C.c_double],
        "psc_b200_mflds_download_interior": [CTX, C.c_int, C.c_int, C.c_int, P],
        "psc_b200_checkpoint_write": [CTX, C.c_char_p, C.c_int64],
        "psc_b200_checkpoint_read": [CTX, C.c_char_p, C.POINTER(C.c_int64)],
        "psc_b200_energies": [CTX, P],
        "psc_b200_last_energies": [CTX, P],
        "psc_b200_step": [CTX, C.POINTER(StepParams)],
        "psc_b200_step_begin": [CTX, C.POINTER(StepParams)],
        "psc_b200_step_end": [CTX],
        "psc_b200_mflds_download_async": [CTX, C.c_int, C.c_int, C.c_int, P],
        "psc_b200_mflds_upload_async": [CTX, C.c_int, C.c_int, C.c_int, P],
        "psc_b200_io_wait": [CTX],
        "psc_b200_last_checks": [CTX, C.POINTER(C.c_double), C.POINTER(C.c_double)],
        "psc_b200_nccl_unique_id": [P],
        "psc_b200_nccl_init": [CTX, P],
        "psc_b200_balance": [CTX, C.c_double, C.POINTER(C.c_int)],
        "psc_b200_best_mapping": [C.c_int, P, C.c_int, P, P],
        "psc_b200_set_option": [CTX, C.c_char_p, C.c_double],
        "psc_b200_get_stat": [CTX, C.c_char_p, C.POINTER(C.c_double)],
        "psc_b200_timer_start": [CTX],
        "psc_b200_timer_stop": [CTX, C.POINTER(C.c_float)],
        "psc_b200_prof_get": [CTX, C.c_int, P, P, P],
        "psc_b200_prof_reset": [CTX],
    }
    for name, argtypes in sigs.items():
        fn = getattr(L, name)  # AttributeError if the .so does not export it
        fn.argtypes = argtypes
        fn.restype = C.c_int
    L.psc_b200_destroy.argtypes = [CTX]
    L.psc_b200_destroy.restype = None
    L._symbols = list(sigs) + ["psc_b200_destroy", "psc_b200_last_error", "psc_b200_version"]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        msg = load().psc_b200_last_error()
        raise PscB200Error(msg.decode() if msg else f"psc_b200 error {rc}")
