"""ctypes binding of the C ABI declared in ``include/qbot_b200.h``.

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a) as
``qbot_b200/lib/libqbot_b200.so``.  There is no fallback: if the library is missing, or no
CUDA device is present, every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libqbot_b200.so')

c_state_p = C.c_void_p


class QbStats(C.Structure):
    _fields_ = [('kernel_launches', C.c_uint64), ('gates_applied', C.c_uint64), ('state_passes', C.c_uint64),
                ('fused_passes', C.c_uint64), ('fused_gates', C.c_uint64), ('bytes_moved', C.c_uint64),
                ('jit_passes', C.c_uint64), ('jit_kernel_hash', C.c_uint64)]


class QbotB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"qbot_b200 [{code}]: {msg}")
        self.code = code


# name -> (restype, argtypes); every entry must exist in include/qbot_b200.h (tests check both ways)
PROTOTYPES = {
    'qb_version': (C.c_char_p, []),
    'qb_last_error': (C.c_char_p, []),
    'qb_device_count': (C.c_int, [C.POINTER(C.c_int)]),
    'qb_device_info': (C.c_int, [C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_size_t),
                                 C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    'qb_create': (C.c_int, [C.POINTER(c_state_p), C.c_int, C.c_int, C.c_int64, C.c_int]),
    'qb_create_external': (C.c_int, [C.POINTER(c_state_p), C.c_int, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    'qb_destroy': (C.c_int, [c_state_p]),
    'qb_clone': (C.c_int, [c_state_p, C.POINTER(c_state_p)]),
    'qb_info': (C.c_int, [c_state_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
    'qb_device_ptr': (C.c_int, [c_state_p, C.POINTER(C.c_void_p)]),
    'qb_set_stream': (C.c_int, [c_state_p, C.c_void_p]),
    'qb_init_basis': (C.c_int, [c_state_p, C.c_uint64]),
    'qb_init_product': (C.c_int, [c_state_p, C.c_void_p, C.c_int]),
    'qb_init_diag': (C.c_int, [c_state_p, C.c_void_p]),
    'qb_upload': (C.c_int, [c_state_p, C.c_void_p, C.c_size_t]),
    'qb_download': (C.c_int, [c_state_p, C.c_void_p, C.c_size_t]),
    'qb_download_range': (C.c_int, [c_state_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    'qb_apply_gate': (C.c_int, [c_state_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_uint64]),
    'qb_apply_gates': (C.c_int, [c_state_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.c_void_p]),
    'qb_apply_swap': (C.c_int, [c_state_p, C.c_int, C.c_int]),
    'qb_apply_gate_batched': (C.c_int, [c_state_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.c_void_p]),
    'qb_apply_gate_rc': (C.c_int, [c_state_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    'qb_flush': (C.c_int, [c_state_p]),
    'qb_sync': (C.c_int, [c_state_p]),
    'qb_set_fusion': (C.c_int, [c_state_p, C.c_int]),
    'qb_set_jit': (C.c_int, [c_state_p, C.c_int]),
    'qb_jit_info': (C.c_int, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_double)]),
    'qb_jit_check': (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.c_void_p,
                               C.POINTER(C.c_int), C.c_char_p]),
    'qb_probs': (C.c_int, [c_state_p, C.POINTER(C.c_int), C.c_int, C.c_void_p]),
    'qb_probs_basis': (C.c_int, [c_state_p, C.POINTER(C.c_int), C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    'qb_norm2': (C.c_int, [c_state_p, C.c_void_p]),
    'qb_project_renorm': (C.c_int, [c_state_p, C.POINTER(C.c_int), C.c_int, C.c_uint64]),
    'qb_ptrace': (C.c_int, [c_state_p, C.POINTER(C.c_int), C.c_int, C.POINTER(c_state_p)]),
    'qb_scatter_product': (C.c_int, [c_state_p, c_state_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.POINTER(c_state_p)]),
    'qb_mix': (C.c_int, [C.POINTER(c_state_p), C.POINTER(C.c_double), C.c_int, C.POINTER(c_state_p)]),
    'qb_mix_branches': (C.c_int, [c_state_p, C.POINTER(C.c_double), C.POINTER(c_state_p)]),
    'qb_outer': (C.c_int, [c_state_p, C.c_int, C.POINTER(c_state_p)]),
    'qb_broadcast': (C.c_int, [c_state_p, c_state_p]),
    'qb_buffer_alloc': (C.c_int, [C.c_int, C.c_size_t, C.POINTER(C.c_void_p)]),
    'qb_buffer_free': (C.c_int, [C.c_int, C.c_void_p]),
    'qb_rebind': (C.c_int, [c_state_p, C.c_void_p]),
    'qb_ipc_export': (C.c_int, [C.c_int, C.c_void_p, C.c_void_p]),
    'qb_ipc_open': (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    'qb_ipc_close': (C.c_int, [C.c_int, C.c_void_p]),
    'qb_permute_scatter': (C.c_int, [c_state_p, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p), C.c_int]),
    'qb_permute_scatter_sub': (C.c_int, [c_state_p, C.POINTER(C.c_int), C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int]),
    'qb_set_sm_limit': (C.c_int, [c_state_p, C.c_int]),
    'qb_plan_queue': (C.c_int, [c_state_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    'qb_run_steps': (C.c_int, [c_state_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    'qb_finish_queue': (C.c_int, [c_state_p]),
    'qb_signal_flags': (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_uint64]),
    'qb_wait_flags': (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_uint64]),
    'qb_flag_timeouts': (C.c_int, [C.c_int, C.POINTER(C.c_uint64)]),
    'qb_compute_stream': (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    'qb_copy_async': (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    'qb_get_stats': (C.c_int, [c_state_p, C.POINTER(QbStats)]),
    'qb_reset_stats': (C.c_int, [c_state_p]),
    'qb_timer_start': (C.c_int, [c_state_p]),
    'qb_timer_stop': (C.c_int, [c_state_p, C.POINTER(C.c_float)]),
}

_lock = threading.Lock()
_lib = None


def _locate_nvrtc():
    """The sweep specialiser dlopens libnvrtc.so.12 by name, then from /usr/local/cuda/lib64; if
    neither exists point it (QBOT_B200_NVRTC) at the copy that ships with the CUDA wheels."""
    if os.environ.get('QBOT_B200_NVRTC'):
        return
    import glob
    import sys
    for pat in ('/usr/local/cuda/lib64/libnvrtc.so.12', '/usr/local/cuda*/lib64/libnvrtc.so.12'):
        if glob.glob(pat):
            return
    for base in sys.path:
        hits = glob.glob(os.path.join(base, 'nvidia', 'cuda_nvrtc', 'lib', 'libnvrtc.so.12'))
        if hits:
            os.environ['QBOT_B200_NVRTC'] = hits[0]
            return


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise QbotB200Error(-2, f"{LIB_PATH} not found -- build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                                    "there is no CPU fallback")
        _locate_nvrtc()
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)        # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(code: int):
    if code != 0:
        raise QbotB200Error(code, load().qb_last_error().decode(errors='replace'))


def call(name: str, *args):
    check(getattr(load(), name)(*args))


def device_count() -> int:
    n = C.c_int(0)
    call('qb_device_count', C.byref(n))
    return n.value


def int_array(vals):
    vals = [int(v) for v in vals]
    return (C.c_int * max(len(vals), 1))(*vals)


def jit_info() -> dict:
    """Process-wide counters of the sweep specialiser."""
    k, h, ms = C.c_uint64(0), C.c_uint64(0), C.c_double(0)
    call('qb_jit_info', C.byref(k), C.byref(h), C.byref(ms))
    return {'kernels_compiled': k.value, 'cache_hits': h.value, 'compile_ms': ms.value}


def jit_check(nbits: int, gate_list, cubin_dir: str = None) -> int:
    """Plan `gate_list` ([(matrix, target_bits msb-first, control_mask)] on index bits), generate
    the specialised source of every fused sweep and NVRTC-compile it for sm_100a without running
    it (needs no GPU).  Returns the number of kernels compiled."""
    import numpy as np
    n = len(gate_list)
    ks = (C.c_int * max(n, 1))()
    tbs = (C.c_int * (14 * max(n, 1)))()
    cms = (C.c_uint64 * max(n, 1))()
    mats = []
    for i, (m, tb, cm) in enumerate(gate_list):
        m = np.ascontiguousarray(np.asarray(m, dtype=np.complex128))
        ks[i] = len(tb)
        for j, b in enumerate(tb):
            tbs[14 * i + j] = int(b)
        cms[i] = int(cm)
        mats.append(m.reshape(-1))
    allm = np.ascontiguousarray(np.concatenate(mats)) if mats else np.zeros(1, dtype=np.complex128)
    out = C.c_int(0)
    call('qb_jit_check', nbits, n, ks, tbs, cms, allm.ctypes.data_as(C.c_void_p), C.byref(out),
         cubin_dir.encode() if cubin_dir else None)
    return out.value
