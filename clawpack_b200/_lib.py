"""ctypes binding of libclawcuda.so (include/clawcuda.h).  There is NO CPU fallback: if the
library cannot be loaded, or no CUDA device is present when a solver is created, the CUDA
path raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('CLAWCUDA_LIB') or os.path.join(_HERE, 'libclawcuda.so')   # override: tuning A/B builds

_lib = None

c_int_p = ctypes.POINTER(ctypes.c_int)
c_double_p = ctypes.POINTER(ctypes.c_double)


class ClawCudaError(RuntimeError):
    pass


def _declare(L):
    i, d, vp, sz = ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_size_t
    H = vp
    sig = {
        'claw_version': (ctypes.c_char_p, []),
        'claw_last_error': (ctypes.c_char_p, [H]),
        'claw_device_count': (i, []),
        'claw_create': (i, [ctypes.POINTER(H), i, i, i, i, i, c_int_p, i, i, vp]),
        'claw_destroy': (i, [H]),
        'claw_buffer_sizes': (i, [H, ctypes.POINTER(sz), ctypes.POINTER(sz), ctypes.POINTER(sz)]),
        'claw_bind': (i, [H, vp, vp, vp, vp]),
        'claw_set_cparam': (i, [H, c_double_p, i]),
        'claw_set_method': (i, [H, c_int_p, c_int_p, i]),
        'claw_set_bc': (i, [H, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p]),
        'claw_set_boundary_ownership': (i, [H, c_int_p, c_int_p]),
        'claw_set_mode': (i, [H, i]),
        'claw_set_source': (i, [H, i, c_double_p, i]),
        'claw_set_bc_state': (i, [H, i, i, c_double_p]),
        'claw_set_tuning': (i, [H, i]),
        'claw_put_q': (i, [H, vp]),
        'claw_get_q': (i, [H, vp]),
        'claw_put_aux': (i, [H, vp]),
        'claw_get_cells': (i, [H, i, c_int_p, c_double_p]),
        'claw_put_qbc': (i, [H, vp]),
        'claw_get_qbc': (i, [H, vp]),
        'claw_put_auxbc': (i, [H, vp]),
        'claw_apply_bcs': (i, [H]),
        'claw_step_launch': (i, [H, c_double_p, d, i]),
        'claw_cfl_device_ptr': (i, [H, ctypes.POINTER(vp)]),
        'claw_cfl_wait': (i, [H, c_double_p]),
        'claw_step': (i, [H, c_double_p, d, i, c_double_p]),
        'claw_accept': (i, [H]),
        'claw_reject': (i, [H]),
        'claw_sync': (i, [H]),
        'claw_ctl_begin': (i, [H, vp, d, i, ctypes.c_longlong]),
        'claw_ctl_attempts': (i, [H, c_double_p, i, i]),
        'claw_ctl_sweep': (i, [H, c_double_p, i]),
        'claw_ctl_end': (i, [H]),
        'claw_ctl_read': (i, [H, vp, ctypes.c_longlong, vp, i, c_int_p]),
        'claw_ctl_use_graph': (i, [H, i]),
        'claw_halo_elems': (i, [H, ctypes.POINTER(sz)]),
        'claw_halo_pack': (i, [H, i, vp]),
        'claw_halo_unpack': (i, [H, i, vp]),
        'claw_halo_pack_next': (i, [H, i, vp]),
        'claw_halo_unpack_next': (i, [H, i, vp]),
        'claw_halo_elems_aux': (i, [H, ctypes.POINTER(sz)]),
        'claw_halo_pack_aux': (i, [H, i, vp]),
        'claw_halo_unpack_aux': (i, [H, i, vp]),
        'claw_halo_exchange': (i, [H, vp, i, i, i, i, i]),
        'claw_allreduce_cfl': (i, [H, vp]),
        'claw_nccl_unique_id': (i, [vp]),
        'claw_nccl_comm_init': (i, [ctypes.POINTER(ctypes.c_void_p), i, i, vp, i]),
        'claw_nccl_comm_destroy': (i, [vp]),
        'claw_cfl_read': (i, [H, c_double_p]),
        'claw_own_buffers': (i, [H]),
        'claw_peer_export': (i, [H, vp]),
        'claw_peer_connect': (i, [H, i, i, vp, i, i]),
        'claw_peer_sync': (i, [H, i]),
        'claw_peer_allreduce_cfl': (i, [H, c_double_p]),
        'claw_time_sweep': (i, [H, c_double_p, d, i, c_double_p]),
        'claw_launch_count': (i, [H, ctypes.POINTER(ctypes.c_uint64)]),
        'claw_last_launch_info': (i, [H, c_int_p]),
        'claw_dense_state': (i, [H, c_int_p]),
        'claw_step2_host': (i, [i] * 7 + [vp, vp, vp, d, d, d, c_int_p, c_int_p, i, i, c_double_p, i, i, c_double_p]),
        'claw_step2ds_host': (i, [i] * 7 + [vp, vp, vp, d, d, d, c_int_p, c_int_p, i, i, i, c_double_p, i, i,
                                            c_double_p]),
        'claw_step3_host': (i, [i] * 8 + [vp, vp, vp, d, d, d, d, c_int_p, c_int_p, i, i, c_double_p, i, i,
                                          c_double_p]),
        'claw_step3ds_host': (i, [i] * 8 + [vp, vp, vp, d, d, d, d, c_int_p, c_int_p, i, i, i, c_double_p, i, i,
                                            c_double_p]),
        'claw_step1_host': (i, [i] * 5 + [vp, vp, d, d, c_int_p, c_int_p, i, i, c_double_p, i, i, c_double_p]),
        'claw_host_release': (i, []),
        'claw_host_alloc': (i, [ctypes.POINTER(vp), sz]),
        'claw_host_free': (i, [vp]),
        'claw_measure_fp64_peak': (i, [i, c_double_p]),
        'claw_rp_plugin_load': (i, [ctypes.c_char_p, c_int_p]),
        'claw_rp_plugin_info': (i, [i, c_int_p]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    return sig


EXPORTS = None


def lib():
    """Load libclawcuda.so (raises ClawCudaError if it was not built)."""
    global _lib, EXPORTS
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ClawCudaError('%s not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                                '(kernel_language="CUDA" has no CPU fallback)' % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        EXPORTS = _declare(L)
        _lib = L
    return _lib


def check(rc, handle=None):
    if rc != 0:
        msg = lib().claw_last_error(handle)
        raise ClawCudaError('libclawcuda error %d: %s' % (rc, msg.decode() if msg else '?'))


def ip(a):
    return a.ctypes.data_as(c_int_p)


def dp(a):
    return a.ctypes.data_as(c_double_p)
