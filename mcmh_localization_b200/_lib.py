"""ctypes binding of libmcl.so (include/mcl.h).  No CPU fallback: if the shared object is missing
or no CUDA device is present, everything here raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmcl.so")

RESAMPLE_REFERENCE_F32 = 0
RESAMPLE_FIXED_POINT = 1
RESAMPLE_AMCL_F32 = 2

STREAM_MOTION, STREAM_MH, STREAM_RESAMPLE, STREAM_INIT, STREAM_KLD = 1, 2, 3, 4, 5

_vp, _i, _i64, _u64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double
_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int)

# name -> (restype, argtypes); the list is also what tests/test_abi.py checks against include/mcl.h
SIGNATURES = {
    "mcl_create": (_i, [C.POINTER(_vp), _i]),
    "mcl_destroy": (_i, [_vp]),
    "mcl_last_error": (C.c_char_p, [_vp]),
    "mcl_version": (C.c_char_p, []),
    "mcl_set_stream": (_i, [_vp, _vp]),
    "mcl_sync": (_i, [_vp]),
    "mcl_device_info": (_i, [_vp, _pi, _pi, _pi, _pi]),
    "mcl_set_map": (_i, [_vp, _vp, _vp, _i, _i, _d, _d, _d]),
    "mcl_set_map_edt": (_i, [_vp, _vp, _i, _i, _d, _d, _d, _vp]),
    "mcl_set_sensor": (_i, [_vp, _d, _d, _d, _d, _i]),
    "mcl_set_motion": (_i, [_vp, C.POINTER(C.c_float)]),
    "mcl_set_scan": (_i, [_vp, _vp, _vp, _i]),
    "mcl_set_scan_batch": (_i, [_vp, _vp, _vp, _i, _i]),
    "mcl_use_scan": (_i, [_vp, _i]),
    "mcl_scan_valid_count": (_i, [_vp, _pi]),
    "mcl_set_likelihood_path": (_i, [_vp, _i]),
    "mcl_likelihood": (_i, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "mcl_set_raycast_grid": (_i, [_vp, _vp, _i, _i, _d, _d, _d]),
    "mcl_likelihood_raycast": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _i64, _vp]),
    "mcl_softmax": (_i, [_vp, _vp, _i64, _vp, _vp, _pd]),
    "mcl_softmax_max": (_i, [_vp, _vp, _i64, _vp]),
    "mcl_softmax_sumexp": (_i, [_vp, _vp, _i64, _vp]),
    "mcl_softmax_weights": (_i, [_vp, _vp, _i64, _vp, _vp]),
    "mcl_weights_normalize": (_i, [_vp, _vp, _i64, _pd]),
    "mcl_softmax_stats": (_i, [_vp, _vp, _i64, _pd]),
    "mcl_predict": (_i, [_vp, _vp, _vp, _vp, _i64, _pd, _u64, _u64, _u64, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "mcl_compute_motion": (_i, [_pd, _pd, _pd]),
    "mcl_mh_accept": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _u64, _u64, _u64,
                           _vp, _vp, _vp, _vp, _vp]),
    "mcl_motion_density": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _pd, _vp, _vp, _i]),
    "mcl_scale_by_sum": (_i, [_vp, _vp, _i64, _vp]),
    "mcl_assym_mh_accept": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _u64, _u64, _u64,
                                 _vp, _vp, _vp, _vp, _vp]),
    "mcl_assym_mh_accept_ex": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _u64, _u64, _u64,
                                 _vp, _vp, _vp, _vp, _vp, _i]),
    "mcl_resample_indices": (_i, [_vp, _vp, _i64, _i64, _d, _i, _vp]),
    "mcl_weights_max": (_i, [_vp, _vp, _i64, _vp]),
    "mcl_resample_scan": (_i, [_vp, _vp, _i64, _vp, _i64, _vp]),
    "mcl_resample_search": (_i, [_vp, _i64, _u64, _u64, _i64, _i64, _d, _i64, _vp]),
    "mcl_kld_resample": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _d, _d, _d, _d, _d, _vp, _u64, _u64, _i,
                              _vp, _vp, _vp, C.POINTER(_i64)]),
    "mcl_resample_push": (_i, [_vp, _i64, _vp, _i, _i, _d, _i64, _i64, _vp, _vp, _vp, _vp]),
    "mcl_resample_offset": (_d, [_u64, _u64, _i64]),
    "mcl_gather": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "mcl_estimate": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _pd]),
    "mcl_estimate_async": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "mcl_estimate_moments_async": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "mcl_estimate_means_async": (_i, [_vp, _vp]),
    "mcl_estimate_central_async": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "mcl_estimate_moments": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _pd]),
    "mcl_estimate_central": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _pd, _pd]),
    "mcl_normalize_angle_array": (_i, [_vp, _vp, _d, _i64, _vp]),
    "mcl_init_uniform": (_i, [_vp, _i64, _vp, _i64, _u64, _u64, _vp, _vp, _vp, C.POINTER(_i64)]),
    "mcl_aos_to_soa": (_i, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "mcl_soa_to_aos": (_i, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "mcl_filter_bind": (_i, [_vp, _i64, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), _vp, _vp, _vp, _vp, _vp, _vp,
                             _vp, _i, _i, _u64, _u64, _i]),
    "mcl_filter_configure": (_i, [_vp, _i, _i, _u64, _u64, _i64]),
    "mcl_filter_set_assym": (_i, [_vp, _i]),
    "mcl_filter_set_transition": (_i, [_vp, _pd, _pd]),
    "mcl_filter_set_n": (_i, [_vp, _i64]),
    "mcl_filter_roles": (_i, [_vp, _pi, C.POINTER(_u64)]),
    "mcl_filter_set_roles": (_i, [_vp, _pi, _u64]),
    "mcl_filter_predict": (_i, [_vp, _pd, _vp, _i]),
    "mcl_filter_update": (_i, [_vp, _vp]),
    "mcl_filter_update_chain": (_i, [_vp, _i]),
    "mcl_filter_estimate": (_i, [_vp, _vp, _pd]),
    "mcl_filter_resample": (_i, [_vp, _d]),
    "mcl_filter_finish": (_i, [_vp, _vp, _vp]),
    "mcl_filter_step": (_i, [_vp, _pd, _i, _vp, _pd]),
    "mcl_comm_init": (_i, [_vp, _i, _i, _vp, C.POINTER(_u64), C.POINTER(_u64)]),
    "mcl_comm_status": (_i, [_vp, _pi]),
    "mcl_tail_status": (_i, [_vp, _pi]),
    "mcl_compute_valid_indices": (_i, [_vp, _vp, _vp, _i64, _vp, C.POINTER(_i64)]),
    "mcl_validate_samples": (_i, [_vp, _vp, _vp, _vp, _i64]),
    "mcl_resample_multinomial": (_i, [_vp, _vp, _i64, _i64, _vp, _u64, _u64, _vp]),
    "mcl_reinitialize_particles": (_i, [_vp, _i64, _vp, _vp, _u64, _u64, _vp, _vp, _vp, C.POINTER(_i64)]),
    "mcl_tail_prof": (_i, [_vp, C.POINTER(C.c_uint64), _pi]),
    "mcl_debug_tail_resample": (_i, [_vp, _vp, _i64, _d, _i, _vp, _vp]),
    "mcl_debug_motion_stats": (_i, [_vp, _vp]),
    "mcl_debug_motion": (_i, [_vp, C.c_ulonglong, _i]),
    "mcl_bench_gather": (_i, [_vp, _i, _i64, _i64, _i, _pd]),
    "mcl_debug_seq_cumsum": (_i, [_vp, _vp, _i64, _i, _i, _vp]),
    "mcl_launch_count": (_i64, [_vp]),
    "mcl_timing_start": (_i, [_vp]),
    "mcl_timing_stop": (_i, [_vp, _pd, C.POINTER(_i64)]),
    "mcl_timing_sets": (_i64, [_vp]),
}


class MclError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libmcl error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load():
    """dlopen libmcl.so and declare every entry point.  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s not found: build it with `python -m mcmh_localization_b200.build` "
                "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)           # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class Handle:
    """Owns one mcl_handle; `call` raises MclError on a negative status."""

    def __init__(self, device=0):
        self.lib = load()
        hp = _vp()
        rc = self.lib.mcl_create(C.byref(hp), int(device))
        if rc != 0:
            raise MclError(rc, self.lib.mcl_last_error(None).decode())
        self.h = hp
        self.device = int(device)

    def call(self, name, *args):
        rc = getattr(self.lib, name)(self.h, *args)
        if rc != 0:
            raise MclError(rc, self.lib.mcl_last_error(self.h).decode())
        return rc

    def close(self):
        if getattr(self, "h", None):
            self.lib.mcl_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
