"""TEST INFRASTRUCTURE: ctypes wrapper of the CPU plan emulator (tests/csrc/plan_emulator.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'build', 'libqb_plan_test.so')
SRCS = [os.path.join(ROOT, 'tests', 'csrc', 'plan_emulator.cpp'), os.path.join(ROOT, 'qbot_b200', 'csrc', 'qb_plan.cpp'),
        os.path.join(ROOT, 'qbot_b200', 'csrc', 'qb_jitgen.cpp')]
DEPS = SRCS + [os.path.join(ROOT, 'qbot_b200', 'csrc', f) for f in ('qb_plan.h', 'qb_tile_ops.h', 'qb_gate.h', 'qb_jit.h')]

_lib = None


def lib():
    global _lib
    if _lib is None:
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in DEPS):
            subprocess.check_call(['g++', '-std=c++17', '-O2', '-shared', '-fPIC', '-o', SO] + SRCS)
        _lib = C.CDLL(SO)
        _lib.qbt_run.restype = C.c_int
        _lib.qbt_last_error.restype = C.c_char_p
    return _lib


def pack_gates(gate_list):
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
    return ks, tbs, cms, allm


def run(nbits, gate_list, psi=None, M=12, merge=True, execute=True):
    """gate_list: [(matrix 2^k x 2^k, target_bits (msb first), control_mask)] on index BITS.
    Returns (psi_out or None, stats dict)."""
    n = len(gate_list)
    ks, tbs, cms, allm = pack_gates(gate_list)
    out = None
    ptr = None
    if execute:
        out = np.ascontiguousarray(np.array(psi, dtype=np.complex128))
        ptr = out.ctypes.data_as(C.c_void_p)
    stats = (C.c_longlong * 7)()
    rc = lib().qbt_run(nbits, n, ks, tbs, cms, allm.ctypes.data_as(C.c_void_p), ptr, M, 1 if merge else 0,
                       1 if execute else 0, stats)
    if rc != 0:
        raise RuntimeError(lib().qbt_last_error().decode())
    keys = ('steps', 'fused_sweeps', 'unfused_steps', 'stages', 'ops', 'max_program_bytes', 'fused_gates')
    return out, dict(zip(keys, [int(x) for x in stats]))


def circuit_to_bits(n, gates):
    """qbot_b200.circuits.Gate list -> emulator gate list (reference qubit q -> index bit n-1-q)."""
    out = []
    for g in gates:
        cm = 0
        for c in g.controls:
            cm |= 1 << (n - 1 - c)
        out.append((g.matrix(), [n - 1 - g.target], cm))
    return out
