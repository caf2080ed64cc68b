"""ctypes binding of libplk_b200.so (C ABI in include/plk.h).  Fails loudly: no CPU fallback exists."""
import ctypes
import os

from . import _build

_LIB = None

c_int, c_ll, c_dbl, vp = ctypes.c_int, ctypes.c_longlong, ctypes.c_double, ctypes.c_void_p

PLK_MAX_PIX_TERMS = 6


class PixProg(ctypes.Structure):
    """plk_pixprog of include/plk.h"""
    _fields_ = [('nterm', c_int), ('a', vp * PLK_MAX_PIX_TERMS), ('b', vp * PLK_MAX_PIX_TERMS), ('scale', c_dbl * PLK_MAX_PIX_TERMS)]


_SIGS = {
    'plk_last_error': (ctypes.c_char_p, []),
    'plk_version': (c_int, []),
    'plk_launch_count': (c_ll, []),
    'plk_plan_create': (c_int, [ctypes.POINTER(vp), c_int, c_int, c_int]),
    'plk_plan_destroy': (c_int, [vp]),
    'plk_plan_device_bytes': (c_ll, [vp]),
    'plk_plan_set_seed_threshold': (c_int, [vp, c_int]),
    'plk_plan_nside': (c_int, [vp]),
    'plk_plan_lmax': (c_int, [vp]),
    'plk_alm2map_dev': (c_int, [vp, c_int, vp, vp, vp, vp, vp, vp, vp]),
    'plk_map2alm_dev': (c_int, [vp, c_int, vp, vp, vp, vp, vp, vp, vp]),
    'plk_alm2map_host': (c_int, [vp, c_int, vp, vp, vp, vp]),
    'plk_map2alm_host': (c_int, [vp, c_int, vp, vp, vp, vp]),
    'plk_legendre_synth_dev': (c_int, [vp, c_int, vp, vp, vp, vp, vp, vp, vp]),
    'plk_legendre_anal_dev': (c_int, [vp, c_int, vp, vp, vp, vp, vp, vp, vp]),
    'plk_ring_synth_dev': (c_int, [vp, vp, vp, vp]),
    'plk_ring_anal_dev': (c_int, [vp, vp, vp, vp]),
    'plk_almxfl_dev': (c_int, [c_int, vp, vp, c_int, vp, vp]),
    'plk_alm_axpy_dev': (c_int, [c_ll, c_dbl, vp, vp, vp, vp]),
    'plk_alm_dot_dev': (c_int, [c_int, c_int, vp, vp, vp, vp]),
    'plk_alm_dot2_dev': (c_int, [c_int, c_int, vp, vp, vp, vp, vp, vp]),
    'plk_alm_dotn_dev': (c_int, [c_int, c_int, c_int, ctypes.POINTER(vp), ctypes.POINTER(vp), vp, vp]),
    'plk_alm_dot_fused_dev': (c_int, [c_int, c_int, c_int, ctypes.POINTER(vp), ctypes.POINTER(vp), vp, vp, c_dbl, vp, vp]),
    'plk_alm_axpy2_dev': (c_int, [c_ll, vp, vp, vp, vp, vp, vp]),
    'plk_map2alm_add_dev': (c_int, [vp, c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    'plk_map2alm_pix_dev': (c_int, [vp, c_int, ctypes.POINTER(PixProg), ctypes.POINTER(PixProg), vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    'plk_alm2cl_dev': (c_int, [c_int, vp, vp, vp, vp]),
    'plk_scalar_ratio_dev': (c_int, [vp, vp, c_dbl, vp, vp]),
    'plk_alm_copy_dev': (c_int, [c_int, vp, c_int, vp, vp]),
    'plk_alm_splice_dev': (c_int, [c_int, vp, c_int, vp, c_int, vp, vp]),
    'plk_alm_lincomb_dev': (c_int, [c_ll, c_dbl, vp, c_dbl, vp, vp, vp]),
    'plk_alm_combine_dev': (c_int, [c_int, c_int, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(c_int), vp, vp]),
    'plk_set_lane': (c_int, [c_int]),
    'plk_get_lane': (c_int, []),
    'plk_alm2rlm_dev': (c_int, [c_int, vp, vp, vp]),
    'plk_alm2rlm_from_dev': (c_int, [c_int, c_int, vp, vp, vp]),
    'plk_alm_splice_xfl_dev': (c_int, [c_int, vp, c_int, vp, vp, c_int, c_int, vp, vp]),
    'plk_rlm2alm_dev': (c_int, [c_int, vp, vp, vp]),
    'plk_dense_matvec_dev': (c_int, [c_int, vp, vp, vp, vp]),
    'plk_dist_partition': (c_int, [c_int, c_int, c_int, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    'plk_dist_create': (c_int, [ctypes.POINTER(vp), vp, c_int, c_int, c_int]),
    'plk_dist_destroy': (c_int, [vp]),
    'plk_dist_export': (c_int, [vp, vp]),
    'plk_dist_import': (c_int, [vp, c_int, vp]),
    'plk_dist_set_peer': (c_int, [vp, c_int, vp, vp]),
    'plk_dist_phase_ptrs': (c_int, [vp, ctypes.POINTER(vp), ctypes.POINTER(vp)]),
    'plk_dist_num_m': (c_int, [vp]),
    'plk_dist_pixel_ranges': (c_int, [vp, ctypes.POINTER(c_ll)]),
    'plk_dist_legendre_synth': (c_int, [vp, c_int, vp, vp, vp, vp, vp]),
    'plk_dist_ring_synth': (c_int, [vp, c_int, vp, vp, vp]),
    'plk_dist_ring_anal': (c_int, [vp, c_int, vp, vp, vp]),
    'plk_dist_legendre_anal': (c_int, [vp, c_int, vp, vp, vp, vp, vp]),
    'plk_dist_legendre_anal_add': (c_int, [vp, c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    'plk_wignerpos_dev': (c_int, [vp, c_int, vp, c_int, c_int, c_int, vp, vp]),
    'plk_wignercoeff_dev': (c_int, [vp, vp, c_int, c_int, c_int, c_int, vp, vp]),
    'plk_profile_enable': (c_int, [c_int]),
    'plk_profile_read': (c_int, [ctypes.POINTER(c_int), ctypes.POINTER(c_dbl)]),
    'plk_plan_active_fraction': (c_int, [vp, c_int, ctypes.POINTER(c_dbl)]),
    'plk_fp64_peak': (c_int, [ctypes.POINTER(c_dbl), c_int]),
    'plk_randn_dev': (c_int, [ctypes.c_ulonglong, ctypes.c_ulonglong, c_ll, c_dbl, vp, vp, vp]),
    'plk_randn_alm_dev': (c_int, [ctypes.c_ulonglong, ctypes.c_ulonglong, c_int, vp, vp]),
    'plk_philox_words_dev': (c_int, [ctypes.c_ulonglong, ctypes.c_ulonglong, c_ll, vp, vp]),
    'plk_map_mul_dev': (c_int, [c_ll, vp, vp, vp]),
    'plk_map_dot_dev': (c_int, [c_ll, vp, vp, vp, vp]),
    'plk_map_mul2_dev': (c_int, [c_ll, vp, vp, vp, vp]),
    'plk_map_qe_pp_dev': (c_int, [c_ll, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    'plk_map_cmul_acc_dev': (c_int, [c_ll, vp, vp, vp, vp, vp, vp, vp]),
    'plk_map_ninv3_dev': (c_int, [c_ll, vp, vp, vp, vp, vp, vp]),
    'plk_udgrade_sum_dev': (c_int, [c_int, vp, c_int, vp, vp]),
    'plk_map_modes_dot_dev': (c_int, [vp, vp, vp, vp, vp]),
    'plk_map_modes_sub_dev': (c_int, [vp, vp, vp, vp, vp, vp]),
}

EXPORTS = tuple(_SIGS.keys())


class PlkError(RuntimeError):
    pass


def load(path=None):
    """dlopen the library and set argtypes; raises if it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = path or os.environ.get('PLK_LIB_PATH') or _build.SO
    if not os.path.exists(path):
        raise PlkError("libplk_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "or `python -m plancklens_b200._build`; there is no CPU fallback." % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)     # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc):
    if rc != 0:
        raise PlkError("libplk_b200 error %d: %s" % (rc, load().plk_last_error().decode()))


def launch_count():
    return int(load().plk_launch_count())
