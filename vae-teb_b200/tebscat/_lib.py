"""ctypes binding of the C ABI declared in include/tebscat.h.

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).
There is no fallback: if the library is missing or no CUDA device is usable the
product path raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('TEBSCAT_LIB', os.path.join(_HERE, 'libtebscat.so'))
ABI_VERSION = 2

TEBSCAT_OK, TEBSCAT_EINVAL, TEBSCAT_ECUDA, TEBSCAT_EUNSUPPORTED = 0, 1, 2, 3


class PlanDesc(ctypes.Structure):
    _fields_ = [('abi_version', ctypes.c_int32), ('N', ctypes.c_int32), ('log2_Np', ctypes.c_int32),
                ('pad_left', ctypes.c_int32), ('n_paths', ctypes.c_int32), ('n_out', ctypes.c_int32),
                ('n_threads', ctypes.c_int32), ('smem_complex', ctypes.c_int32),
                ('n_tasks', ctypes.c_int32), ('n_steps', ctypes.c_int32),
                ('border_mode', ctypes.c_int32), ('scratch_complex', ctypes.c_int32),
                ('reserved', ctypes.c_int32 * 4)]


class PhaseDesc(ctypes.Structure):
    _fields_ = [('abi_version', ctypes.c_int32), ('N', ctypes.c_int32), ('n_filters', ctypes.c_int32),
                ('n_pairs', ctypes.c_int32), ('n_out', ctypes.c_int32), ('n_cols_pad', ctypes.c_int32),
                ('reserved', ctypes.c_int32 * 10)]


class Epilogue(ctypes.Structure):
    _fields_ = [('mean_dev', ctypes.c_void_p), ('std_dev', ctypes.c_void_p), ('mode_dev', ctypes.c_void_p),
                ('log_eps', ctypes.c_float), ('trim', ctypes.c_int32), ('time_major', ctypes.c_int32)]


_lib = None


class TebscatError(RuntimeError):
    pass


def load():
    """Load libtebscat.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TebscatError(
            'tebscat CUDA library not found at %s; build it with '
            '`python -c "import __graft_entry__ as g; g.build()"` from the repository root. '
            'There is no CPU fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32p, fp = ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_float)
    lib.tebscat_abi_version.restype = ctypes.c_int
    lib.tebscat_last_error.restype = ctypes.c_char_p
    lib.tebscat_last_launch_count.restype = ctypes.c_int
    lib.tebscat_plan_create.restype = ctypes.c_int
    lib.tebscat_plan_create.argtypes = [ctypes.POINTER(PlanDesc), fp, ctypes.c_size_t, i32p, i32p,
                                        i32p, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(vp)]
    lib.tebscat_plan_save.restype = ctypes.c_int
    lib.tebscat_plan_save.argtypes = [ctypes.c_char_p, ctypes.POINTER(PlanDesc), fp, ctypes.c_size_t, i32p, i32p, i32p,
                                      ctypes.c_size_t]
    lib.tebscat_plan_get_desc.restype = ctypes.c_int
    lib.tebscat_plan_get_desc.argtypes = [vp, ctypes.POINTER(PlanDesc)]
    lib.tebscat_plan_load.restype = ctypes.c_int
    lib.tebscat_plan_load.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(vp)]
    lib.tebscat_plan_set_window.restype = ctypes.c_int
    lib.tebscat_plan_set_window.argtypes = [vp, fp]
    lib.tebscat_phase_plan_set_window.restype = ctypes.c_int
    lib.tebscat_phase_plan_set_window.argtypes = [vp, fp]
    lib.tebscat_large_set_window.restype = ctypes.c_int
    lib.tebscat_large_set_window.argtypes = [vp, fp, ctypes.c_int]
    lib.tebscat_plan_destroy.restype = None
    lib.tebscat_plan_destroy.argtypes = [vp]
    lib.tebscat_scat1d_forward.restype = ctypes.c_int
    lib.tebscat_scat1d_forward.argtypes = [vp, vp, ctypes.c_int64, vp, vp]
    lib.tebscat_scat1d_forward_gsrc.restype = ctypes.c_int
    lib.tebscat_scat1d_forward_gsrc.argtypes = [vp, vp, ctypes.c_int64, ctypes.c_int64, vp, vp]
    lib.tebscat_scat1d_forward_ex.restype = ctypes.c_int
    lib.tebscat_scat1d_forward_ex.argtypes = [vp, vp, ctypes.c_int64, vp, ctypes.POINTER(Epilogue), vp]
    lib.tebscat_scat1d_forward_host.restype = ctypes.c_int
    lib.tebscat_scat1d_forward_host.argtypes = [vp, vp, ctypes.c_int64, vp]
    lib.tebscat_scat1d_host_copies_only.restype = ctypes.c_int
    lib.tebscat_scat1d_host_copies_only.argtypes = [vp, vp, ctypes.c_int64, vp]
    lib.tebscat_phase_plan_create.restype = ctypes.c_int
    lib.tebscat_phase_plan_create.argtypes = [ctypes.POINTER(PhaseDesc), vp, fp, i32p, i32p, fp, ctypes.POINTER(vp)]
    lib.tebscat_phase_plan_create_pairs_only.restype = ctypes.c_int
    lib.tebscat_phase_plan_create_pairs_only.argtypes = [ctypes.POINTER(PhaseDesc), ctypes.c_int, fp, i32p, i32p, fp, ctypes.POINTER(vp)]
    lib.tebscat_phase_pairs.restype = ctypes.c_int
    lib.tebscat_phase_pairs.argtypes = [vp, vp, vp, ctypes.c_int64, i32p, ctypes.c_int, ctypes.c_int, vp, vp]
    lib.tebscat_large_storez.restype = ctypes.c_int
    lib.tebscat_large_storez.argtypes = [vp, vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_int, vp, vp, vp]
    lib.tebscat_phase_plan_profile.restype = ctypes.c_int
    lib.tebscat_phase_plan_profile.argtypes = [vp, ctypes.c_int]
    lib.tebscat_phase_plan_profile_read.restype = ctypes.c_int
    lib.tebscat_phase_plan_profile_read.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                                    ctypes.POINTER(ctypes.c_int)]
    lib.tebscat_phase_plan_attach_pair_plan.restype = ctypes.c_int
    lib.tebscat_phase_plan_attach_pair_plan.argtypes = [vp, vp]
    lib.tebscat_phase_plan_destroy.restype = None
    lib.tebscat_phase_plan_destroy.argtypes = [vp]
    lib.tebscat_phase_forward.restype = ctypes.c_int
    lib.tebscat_phase_forward.argtypes = [vp, vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, i32p,
                                          ctypes.c_int, ctypes.c_int, vp, vp]
    lib.tebscat_phase_forward_dual.restype = ctypes.c_int
    lib.tebscat_phase_forward_dual.argtypes = [vp, vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               i32p, ctypes.c_int, i32p, ctypes.c_int, vp, vp, vp]
    u32 = ctypes.c_uint32
    lib.tebscat_large_create.restype = ctypes.c_int
    lib.tebscat_large_create.argtypes = [ctypes.c_int, ctypes.POINTER(vp)]
    lib.tebscat_large_destroy.restype = None
    lib.tebscat_large_destroy.argtypes = [vp]
    lib.tebscat_large_set_tile_plan.restype = ctypes.c_int
    lib.tebscat_large_set_tile_plan.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp]
    lib.tebscat_large_pad_load.restype = ctypes.c_int
    lib.tebscat_large_pad_load.argtypes = [vp, vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp]
    lib.tebscat_large_pad_load_mode.restype = ctypes.c_int
    lib.tebscat_large_pad_load_mode.argtypes = [vp, vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp]
    lib.tebscat_large_fft.restype = ctypes.c_int
    lib.tebscat_large_fft.argtypes = [vp, vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, vp]
    lib.tebscat_large_pair.restype = ctypes.c_int
    lib.tebscat_large_pair.argtypes = [vp, vp, ctypes.c_int64, ctypes.c_int, vp]
    lib.tebscat_large_mulfold.restype = ctypes.c_int
    lib.tebscat_large_mulfold.argtypes = [vp, vp, vp, vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, u32, ctypes.c_int,
                                          ctypes.c_int, vp]
    lib.tebscat_large_modulus.restype = ctypes.c_int
    lib.tebscat_large_modulus.argtypes = [vp, vp, ctypes.c_int64, vp]
    lib.tebscat_large_store.restype = ctypes.c_int
    lib.tebscat_large_store.argtypes = [vp, vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, vp, vp]
    lib.tebscat_large_leaf.restype = ctypes.c_int
    lib.tebscat_large_leaf.argtypes = [vp, vp, vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, u32, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp]
    lib.tebscat_large_leaf_adjoint.restype = ctypes.c_int
    lib.tebscat_large_leaf_adjoint.argtypes = [vp, vp, vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, u32, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int, vp]
    lib.tebscat_large_unstore_row.restype = ctypes.c_int
    lib.tebscat_large_unstore_row.argtypes = [vp, vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_int, ctypes.c_int, vp, vp]
    lib.tebscat_large_modulus_to.restype = ctypes.c_int
    lib.tebscat_large_modulus_to.argtypes = [vp, vp, vp, ctypes.c_int64, vp]
    lib.tebscat_large_modulus_backward.restype = ctypes.c_int
    lib.tebscat_large_modulus_backward.argtypes = [vp, vp, vp, ctypes.c_int64, vp]
    lib.tebscat_large_unfold.restype = ctypes.c_int
    lib.tebscat_large_unfold.argtypes = [vp, vp, vp, vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, u32, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_int, vp]
    lib.tebscat_large_unstore.restype = ctypes.c_int
    lib.tebscat_large_unstore.argtypes = [vp, vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_int, vp, vp]
    lib.tebscat_large_pad_adjoint.restype = ctypes.c_int
    lib.tebscat_large_pad_adjoint.argtypes = [vp, vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp]
    lib.tebscat_scat1d_profile_steps.restype = ctypes.c_int
    lib.tebscat_scat1d_profile_steps.argtypes = [vp, vp, ctypes.c_int64, vp, vp, vp]
    lib.tebscat_bench_fp32_peak.restype = ctypes.c_int
    lib.tebscat_bench_fp32_peak.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    if lib.tebscat_abi_version() != ABI_VERSION:
        raise TebscatError('libtebscat.so ABI %d != binding %d; rebuild' % (lib.tebscat_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(rc):
    if rc != TEBSCAT_OK:
        msg = load().tebscat_last_error().decode('utf-8', 'replace')
        if rc == TEBSCAT_EUNSUPPORTED:
            raise NotImplementedError(msg)
        if rc == TEBSCAT_EINVAL:
            raise ValueError(msg)
        raise TebscatError(msg)
