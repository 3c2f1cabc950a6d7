"""TEST INFRASTRUCTURE: CPU execution of the sweep specialiser's generated source.

The planner's sweep programs are turned into source text by qj_generate (qbot_b200/csrc/
qb_jitgen.cpp) -- the same text NVRTC compiles into the sm_100a kernel -- and compiled here
with g++ over CPU definitions of its macros (tests/csrc/jit_cpu_prelude.h), so that the code
generator is checked against the oracle without a GPU."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

import plan_emu

ROOT = plan_emu.ROOT
GEN_DIR = os.path.join(ROOT, 'build', 'jit_emu')


def plan(nbits, gate_list, M=12, merge=True):
    """-> list of (fused, gate_index, program bytes or None)"""
    lib = plan_emu.lib()
    n = len(gate_list)
    ks, tbs, cms, allm = plan_emu.pack_gates(gate_list)
    max_steps = 4 * n + 8
    nsteps = C.c_int(0)
    fused = (C.c_int * max_steps)()
    gidx = (C.c_int * max_steps)()
    off = (C.c_longlong * max_steps)()
    ln = (C.c_longlong * max_steps)()
    cap = 8192 * max_steps
    buf = (C.c_ubyte * cap)()
    rc = lib.qbt_plan(nbits, n, ks, tbs, cms, allm.ctypes.data_as(C.c_void_p), M, 1 if merge else 0, max_steps, C.byref(nsteps),
                      fused, gidx, off, ln, buf, C.c_longlong(cap))
    if rc != 0:
        raise RuntimeError(lib.qbt_last_error().decode())
    raw = bytes(buf)
    return [(bool(fused[i]), int(gidx[i]), raw[off[i]:off[i] + ln[i]] if fused[i] else None) for i in range(nsteps.value)]


def source_of(program: bytes) -> str:
    lib = plan_emu.lib()
    lib.qbt_jit_source.restype = C.c_char_p
    return lib.qbt_jit_source(program).decode()


def pool_of(program: bytes) -> np.ndarray:
    lib = plan_emu.lib()
    out = np.zeros(4096, dtype=np.float64)
    n = lib.qbt_jit_pool(program, out.ctypes.data_as(C.c_void_p), out.size)
    assert n > 0
    return out[:n].copy()


def compile_steps(sources):
    """g++ the generated sources (one translation unit each) into one shared object."""
    os.makedirs(GEN_DIR, exist_ok=True)
    key = hashlib.sha1('\n@@\n'.join(sources).encode()).hexdigest()[:16]
    so = os.path.join(GEN_DIR, f'steps_{key}.so')
    if not os.path.exists(so):
        files = []
        for i, src in enumerate(sources):
            f = os.path.join(GEN_DIR, f'steps_{key}_{i}.cpp')
            with open(f, 'w') as fh:
                fh.write('#include "jit_cpu_prelude.h"\n' + src + f'\nQJ_CPU_HARNESS(qj_run_step_{i})\n')
            files.append(f)
        subprocess.check_call(['g++', '-std=c++17', '-O1', '-shared', '-fPIC', '-I', os.path.join(ROOT, 'tests', 'csrc'), '-o', so] + files)
        for f in files:
            os.remove(f)
    return C.CDLL(so)


def run(nbits, gate_list, psi, M=12, merge=True):
    """Apply gate_list through plan -> generated source -> g++ -> execution.  Every step must be
    a fused sweep (the unfused fallbacks are covered by tests/test_planner.py)."""
    steps = plan(nbits, gate_list, M, merge)
    assert all(f for f, _, _ in steps), "circuit contains steps the tile kernel does not run"
    progs = [p for _, _, p in steps]
    sources = [source_of(p) for p in progs]
    lib = compile_steps(sources)
    out = np.ascontiguousarray(np.array(psi, dtype=np.complex128))
    for i, p in enumerate(progs):
        pool = pool_of(p)
        getattr(lib, f'qj_run_step_{i}')(out.ctypes.data_as(C.c_void_p), nbits, pool.ctypes.data_as(C.c_void_p))
    return out, {'sweeps': len(progs), 'source_bytes': sum(len(s) for s in sources)}
