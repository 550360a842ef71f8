"""ctypes declarations of the C ABI in include/eonb200.h.

`bind(cdll)` attaches argument / result types to a loaded library.  The product loader
(`eon_b200._lib`) only ever hands it libeonb200.so; the CPU-only test-suite reuses the same
declarations for its emulation build of the device sources (tests/emu), which is test
infrastructure and never reachable from this package.
"""
from __future__ import annotations

import ctypes as C

POT_LJ, POT_LJCLUSTER, POT_MORSE_PT, POT_EAM = 0, 1, 2, 3
OPT_CG, OPT_LBFGS, OPT_FIRE = 0, 1, 2
MINMODE_DIMER, MINMODE_LANCZOS = 0, 1
ROT_SD, ROT_CG, ROT_LBFGS = 0, 1, 2
METRIC_NORM, METRIC_MAX_ATOM, METRIC_MAX_COMPONENT = 0, 1, 2

ABI_VERSION = 6

#: every symbol include/eonb200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "eonb200_config_default", "eonb200_create", "eonb200_destroy", "eonb200_abi_version",
    "eonb200_last_error", "eonb200_force", "eonb200_force_batch", "eonb200_force_batch_device",
    "eonb200_force_calls", "eonb200_neighbors", "eonb200_search_params_default",
    "eonb200_walkers_create", "eonb200_walkers_destroy", "eonb200_walkers_set",
    "eonb200_walkers_matter_forces", "eonb200_walkers_saddle_search", "eonb200_walkers_relax",
    "eonb200_walkers_displace_along_mode", "eonb200_walkers_harvest", "eonb200_walkers_last_ms",
    "eonb200_walkers_displace", "eonb200_walkers_set_active", "eonb200_walkers_groups", "eonb200_walkers_submit", "eonb200_walkers_poll",
    "eonb200_walkers_harvest_group",
    "eonb200_walkers_queue",
    "eonb200_measure_fp64_peak", "eonb200_kernel_launches", "eonb200_nl_stats",
    "eonb200_last_force_ms", "eonb200_phase_clocks", "eonb200_l2_flush",
    "eonb200_neb_params_default", "eonb200_neb_forces", "eonb200_neb_forces_device", "eonb200_neb_band_forces",
]


class EamRow(C.Structure):
    _fields_ = [("Z", C.c_int), ("Dm", C.c_double), ("alphaM", C.c_double), ("Rm", C.c_double),
                ("beta1", C.c_double), ("beta2", C.c_double), ("r_cut", C.c_double),
                ("func_coeff", C.c_double * 9)]


class NebParams(C.Structure):
    _fields_ = [("spring_constant", C.c_double), ("k_min", C.c_double), ("k_max", C.c_double),
                ("climbing_image", C.c_int), ("old_tangent", C.c_int), ("projection", C.c_int),
                ("energy_weighted", C.c_int)]


NEB_PROJECTION_NEB, NEB_PROJECTION_DNEB, NEB_PROJECTION_ELASTIC_BAND = 0, 1, 2


class Config(C.Structure):
    _fields_ = [("kind", C.c_int), ("device", C.c_int), ("cutoff", C.c_double),
                ("lj_u0", C.c_double), ("lj_psi", C.c_double),
                ("morse_De", C.c_double), ("morse_a", C.c_double), ("morse_re", C.c_double),
                ("n_eam_rows", C.c_int), ("eam_rows", C.POINTER(EamRow)),
                ("eam_conservative", C.c_int), ("skin", C.c_double), ("threads_per_walker", C.c_int),
                ("tiles_override", C.c_int), ("lpa_override", C.c_int),
                ("pin_host_buffers", C.c_int), ("eam_pair_mode", C.c_int)]


class SearchParams(C.Structure):
    _fields_ = [("finite_difference", C.c_double), ("remove_net_force", C.c_int), ("opt_method", C.c_int),
                ("convergence_metric", C.c_int), ("opt_max_iterations", C.c_long), ("max_move", C.c_double),
                ("opt_converged_force", C.c_double), ("time_step", C.c_double), ("max_time_step", C.c_double),
                ("lbfgs_memory", C.c_long), ("lbfgs_inverse_curvature", C.c_double),
                ("lbfgs_auto_scale", C.c_int), ("lbfgs_angle_reset", C.c_int), ("lbfgs_distance_reset", C.c_int),
                ("cg_no_overshooting", C.c_int), ("cg_knock_out_max_move", C.c_int),
                ("cg_max_iter_before_reset", C.c_long), ("dimer_opt_method", C.c_int),
                ("dimer_converged_angle", C.c_double), ("dimer_rotations_max", C.c_long),
                ("lanczos_tolerance", C.c_double), ("lanczos_max_iterations", C.c_long),
                ("lanczos_quit_early", C.c_int), ("minmode_method", C.c_int),
                ("saddle_max_iterations", C.c_long), ("saddle_converged_force", C.c_double),
                ("saddle_max_energy", C.c_double), ("saddle_perp_force_ratio", C.c_double),
                ("saddle_zero_mode_abort_curvature", C.c_double), ("saddle_nonlocal_count_abort", C.c_long),
                ("saddle_nonlocal_distance_abort", C.c_double)]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_long)
_vp = C.c_void_p


def bind(lib: C.CDLL) -> C.CDLL:
    lib.eonb200_config_default.argtypes = [C.POINTER(Config), C.c_int]
    lib.eonb200_config_default.restype = None
    lib.eonb200_create.argtypes = [C.POINTER(Config), C.c_char_p, C.c_size_t]
    lib.eonb200_create.restype = _vp
    lib.eonb200_destroy.argtypes = [_vp]
    lib.eonb200_destroy.restype = None
    lib.eonb200_abi_version.argtypes = []
    lib.eonb200_abi_version.restype = C.c_int
    lib.eonb200_last_error.argtypes = [_vp]
    lib.eonb200_last_error.restype = C.c_char_p
    lib.eonb200_force.argtypes = [_vp, C.c_long, _dp, _ip, _dp, _dp, _dp, _dp]
    lib.eonb200_force.restype = C.c_int
    lib.eonb200_force_batch.argtypes = [_vp, C.c_long, C.c_long, C.POINTER(_dp), C.POINTER(_ip), C.POINTER(_dp),
                                        _dp, _dp, C.POINTER(_dp)]
    lib.eonb200_force_batch.restype = C.c_int
    lib.eonb200_force_batch_device.argtypes = [_vp, C.c_long, C.c_long, _vp, _vp, _vp, _vp, _vp, C.c_int]
    lib.eonb200_force_batch_device.restype = C.c_int
    lib.eonb200_force_calls.argtypes = [_vp]
    lib.eonb200_force_calls.restype = C.c_ulonglong
    lib.eonb200_neighbors.argtypes = [_vp, C.c_long, C.c_long, _dp, _dp, _ip, C.c_double, C.c_int, C.c_long, _lp,
                                      C.POINTER(C.c_int64), C.POINTER(C.c_int32), _dp, _dp, _lp]
    lib.eonb200_neighbors.restype = C.c_int
    lib.eonb200_search_params_default.argtypes = [C.POINTER(SearchParams)]
    lib.eonb200_search_params_default.restype = None
    lib.eonb200_walkers_create.argtypes = [_vp, C.c_long, C.c_long, _ip, _ip, _dp, _dp, _dp, C.c_char_p, C.c_size_t]
    lib.eonb200_walkers_create.restype = _vp
    lib.eonb200_walkers_destroy.argtypes = [_vp]
    lib.eonb200_walkers_destroy.restype = None
    lib.eonb200_walkers_set.argtypes = [_vp, _dp, _dp]
    lib.eonb200_walkers_set.restype = C.c_int
    lib.eonb200_walkers_matter_forces.argtypes = [_vp, C.c_int, _dp, _dp]
    lib.eonb200_walkers_matter_forces.restype = C.c_int
    lib.eonb200_walkers_saddle_search.argtypes = [_vp, C.POINTER(SearchParams), _dp]
    lib.eonb200_walkers_saddle_search.restype = C.c_int
    lib.eonb200_walkers_relax.argtypes = [_vp, C.POINTER(SearchParams)]
    lib.eonb200_walkers_relax.restype = C.c_int
    lib.eonb200_walkers_displace_along_mode.argtypes = [_vp, C.c_double]
    lib.eonb200_walkers_displace_along_mode.restype = C.c_int
    lib.eonb200_walkers_harvest.argtypes = [_vp, _ip, _lp, _lp, _dp, _dp, _dp, _dp, _dp]
    lib.eonb200_walkers_harvest.restype = C.c_int
    lib.eonb200_walkers_displace.argtypes = [_vp, _dp, C.c_long, _ip, C.c_double, C.c_double, C.c_ulonglong, C.c_long]
    lib.eonb200_walkers_displace.restype = C.c_int
    lib.eonb200_walkers_set_active.argtypes = [_vp, _ip]
    lib.eonb200_walkers_set_active.restype = C.c_int
    lib.eonb200_walkers_groups.argtypes = [_vp, C.c_int]
    lib.eonb200_walkers_groups.restype = C.c_int
    lib.eonb200_walkers_submit.argtypes = [_vp, C.c_int, C.c_int, C.POINTER(SearchParams), C.c_long, C.c_long, _dp, _dp, _dp]
    lib.eonb200_walkers_submit.restype = C.c_int
    lib.eonb200_walkers_poll.argtypes = [_vp, C.c_int, C.c_int]
    lib.eonb200_walkers_poll.restype = C.c_int
    lib.eonb200_walkers_harvest_group.argtypes = [_vp, C.c_int, _ip, _lp, _lp, _dp, _dp, _dp, _dp, _dp]
    lib.eonb200_walkers_harvest_group.restype = C.c_int
    lib.eonb200_walkers_queue.argtypes = [_vp, C.c_int, C.POINTER(SearchParams), C.c_long, _dp, _dp, _dp, _ip, _lp, _lp,
                                          _dp, _dp, _dp]
    lib.eonb200_walkers_queue.restype = C.c_int
    lib.eonb200_walkers_last_ms.argtypes = [_vp]
    lib.eonb200_walkers_last_ms.restype = C.c_double
    lib.eonb200_measure_fp64_peak.argtypes = [_vp, C.c_int]
    lib.eonb200_measure_fp64_peak.restype = C.c_double
    lib.eonb200_kernel_launches.argtypes = [_vp]
    lib.eonb200_kernel_launches.restype = C.c_ulonglong
    lib.eonb200_nl_stats.argtypes = [_vp, C.POINTER(C.c_longlong)]
    lib.eonb200_nl_stats.restype = C.c_int
    lib.eonb200_last_force_ms.argtypes = [_vp, _dp]
    lib.eonb200_last_force_ms.restype = C.c_int
    lib.eonb200_phase_clocks.argtypes = [_vp, C.c_int, C.POINTER(C.c_longlong), C.c_long, C.POINTER(C.c_long)]
    lib.eonb200_phase_clocks.restype = C.c_int
    lib.eonb200_l2_flush.argtypes = [_vp, C.c_size_t]
    lib.eonb200_l2_flush.restype = C.c_int
    lib.eonb200_neb_params_default.argtypes = [C.POINTER(NebParams)]
    lib.eonb200_neb_params_default.restype = None
    lib.eonb200_neb_forces.argtypes = [_vp, C.c_long, C.c_long, C.c_long, C.c_long, _dp, _dp, _dp, _dp,
                                       C.POINTER(NebParams), _dp, _dp, C.POINTER(C.c_int)]
    lib.eonb200_neb_forces.restype = C.c_int
    lib.eonb200_neb_forces_device.argtypes = [_vp, C.c_long, C.c_long, C.c_long, C.c_long, C.c_void_p, C.c_void_p,
                                              C.c_void_p, _dp, C.POINTER(NebParams), C.c_void_p, C.c_void_p, C.c_void_p]
    lib.eonb200_neb_forces_device.restype = C.c_int
    lib.eonb200_neb_band_forces.argtypes = [_vp, C.c_long, C.c_long, C.c_long, _dp, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                            _dp, C.c_int, C.POINTER(NebParams), _dp, _dp, _dp, C.POINTER(C.c_int)]
    lib.eonb200_neb_band_forces.restype = C.c_int
    return lib
