"""CPU: the sweep specialiser (qbot_b200/csrc/qb_jitgen.cpp).  The source text it generates for
the planner's sweep programs -- the text NVRTC compiles into the sm_100a kernel -- is compiled
with g++ over CPU definitions of its macros and executed against the oracle: register renaming
of X / CNOT / Toffoli, constant tile addressing, predicates on thread / tile bits, merged phase
tables, controlled diagonals, two-qubit blocks.  One test also NVRTC-compiles the real CUDA
text for sm_100a (no GPU needed for that)."""
import numpy as np
import pytest

import jit_emu
import plan_emu
from oracle import qbot_oracle as orc
from qbot_b200.circuits import rc
from conftest import close
from test_planner import random_gate_list, oracle_apply_bits, rand_ket, rand_u


def tileable(gl):
    """the generated kernels only ever see gates the planner fuses"""
    out = []
    for m, tb, cm in gl:
        m = np.asarray(m)
        diag = np.count_nonzero(m - np.diag(np.diagonal(m))) == 0
        if (diag and len(tb) <= 3) or (not diag and len(tb) <= 2):
            out.append((m, tb, cm))
    return out


@pytest.fixture(params=[4, 5], ids=['R4', 'R5'])
def plan_R(request, monkeypatch):
    """register bits per stage of the plans: 4 (16 amplitudes per thread, also run by the generic
    kernel) and 5 (32 per thread, the specialiser's own plan shape)"""
    monkeypatch.setenv('QBOT_B200_PLAN_R', str(request.param))
    return request.param


@pytest.mark.parametrize('M', [11, 12])
def test_generated_source_rc(M, plan_R):
    for n, depth, seed in ((12, 10, 1), (14, 12, 2), (16, 5, 3)):
        gates = rc(n, depth, seed)
        psi = rand_ket(np.random.default_rng(seed), n)
        out, info = jit_emu.run(n, plan_emu.circuit_to_bits(n, gates), psi, M=M)
        ref = psi
        for g in gates:
            ref = orc.ket_apply(ref, n, g.target, g.matrix(), g.controls)
        assert close(out, ref, 1e-12), (n, info)


@pytest.mark.parametrize('M,merge', [(12, True), (11, True), (12, False)])
def test_generated_source_mixed(M, merge, plan_R):
    rng = np.random.default_rng(300 + M)
    for n in (12, 13, 15):
        gl = tileable(random_gate_list(rng, n, 70))
        psi = rand_ket(rng, n)
        out, info = jit_emu.run(n, gl, psi, M=M, merge=merge)
        ref = psi
        for m, tb, cm in gl:
            ref = oracle_apply_bits(ref, n, m, tb, cm)
        assert close(out, ref, 1e-12), (n, info)


def _flip_heavy(rng, n, count):
    """X / H / diagonal / 2x2 traffic under 0-2 controls: with n close to the tile size most controls are
    thread bits, which is what the deferred-flip rules of the generator (qb_jitgen.cpp, `Pend`) see"""
    X = np.array([[0, 1], [1, 0]], dtype=complex)
    H = np.array([[1, 1], [1, -1]], dtype=complex) * 2 ** -0.5
    Z = np.diag([1, -1]).astype(complex)
    out = []
    for _ in range(count):
        bits = [int(b) for b in rng.permutation(n)]
        r = rng.random()
        nc = int(rng.integers(0, 3))
        if r < 0.40:
            m = X
            nc = max(nc, 1) if rng.random() < 0.8 else nc
        elif r < 0.60:
            m = H
            nc = 0 if rng.random() < 0.8 else nc
        elif r < 0.72:
            m = np.diag(np.exp(1j * rng.uniform(0, 6.28, 2)))
        elif r < 0.80:
            m = Z
        elif r < 0.88:
            m = rand_u(rng, 1)
            nc = min(nc, 1)
        elif r < 0.94:
            out.append((rand_u(rng, 2), bits[:2], 0))
            continue
        else:
            out.append((np.diag(np.exp(1j * rng.uniform(0, 6.28, 4))), bits[:2], 0))
            continue
        cm = 0
        for c in bits[1:1 + nc]:
            cm |= 1 << c
        out.append((m, bits[:1], cm))
    return out


def _cx_then_ch(rng, n, count):
    """controlled X directly followed by a controlled Hadamard (a conditional 2x2) on the same target"""
    X = np.array([[0, 1], [1, 0]], dtype=complex)
    H = np.array([[1, 1], [1, -1]], dtype=complex) * 2 ** -0.5
    out = []
    for _ in range(count):
        b = [int(x) for x in rng.permutation(n)]
        out.append((X, b[:1], 1 << b[1]))
        out.append((H, b[:1], (1 << b[2]) if rng.random() < 0.7 else (1 << b[2]) | (1 << b[3])))
        if rng.random() < 0.3:
            out.append((H, [b[4]], 0))
    return out


def test_deferred_conditional_x(plan_R):
    """A controlled X whose controls are thread / tile bits is not executed as a predicated exchange: it stays a
    pending flip that the store (address choice), a Hadamard (sign), a diagonal (factor exchange), a 2x2
    (column exchange) or a CNOT controlled by its bit (second flip) absorb.  Every rule must occur in the
    generated text of these circuits and every result must match the oracle."""
    import re
    seen = set()
    cases = [(_flip_heavy, seed, 60) for seed in (1, 3, 13, 21)] + [(_cx_then_ch, seed, 20) for seed in (0, 3)]
    for make, seed, count in cases:
        rng = np.random.default_rng(seed)
        n = int(rng.integers(12, 15))
        M = 12 if seed % 3 else 11
        gl = make(rng, n, count)
        psi = rand_ket(rng, n)
        for fused, _, prog in jit_emu.plan(n, gl, M, True):
            assert fused
            seen.update(re.findall(r'\[flip:([a-z0-9-]+)\]', jit_emu.source_of(prog)))
        out, info = jit_emu.run(n, gl, psi, M=M)
        ref = psi
        for m, tb, cm in gl:
            ref = oracle_apply_bits(ref, n, m, tb, cm)
        assert close(out, ref, 1e-12), (make.__name__, seed, n, info)
    assert {'store', 'store-partial', 'h', 'phase', 'cdiag', 'u2', 'u2cond', 'spawn', 'first', 'flush'} <= seen, seen


def test_eager_x_switch(monkeypatch):
    """QBOT_B200_EAGER_X=1 restores the predicated exchanges (A/B timing): same results, no deferred flips"""
    import re
    monkeypatch.setenv('QBOT_B200_EAGER_X', '1')
    rng = np.random.default_rng(13)
    n = 13
    gl = _flip_heavy(rng, n, 50)
    psi = rand_ket(rng, n)
    for _, _, prog in jit_emu.plan(n, gl, 12, True):
        assert not re.findall(r'\[flip:', jit_emu.source_of(prog))
    out, _ = jit_emu.run(n, gl, psi, M=12)
    ref = psi
    for m, tb, cm in gl:
        ref = oracle_apply_bits(ref, n, m, tb, cm)
    assert close(out, ref, 1e-12)


def _sign_heavy(rng, n, count):
    """CZ / CCZ / CNOT / H / RZ traffic: conditional sign flips on thread and tile bits, signs controlled by
    register bits that have a pending exchange, merged phases on the bit of a pending exchange"""
    X = np.array([[0, 1], [1, 0]], dtype=complex)
    H = np.array([[1, 1], [1, -1]], dtype=complex) * 2 ** -0.5
    Z = np.diag([1, -1]).astype(complex)
    out = []
    for _ in range(count):
        b = [int(x) for x in rng.permutation(n)]
        r = rng.random()
        if r < 0.30:
            out.append((Z, b[:1], (1 << b[1]) | ((1 << b[2]) if rng.random() < 0.4 else 0)))
        elif r < 0.60:
            out.append((X, b[:1], (1 << b[1]) | ((1 << b[2]) if rng.random() < 0.3 else 0)))
        elif r < 0.80:
            out.append((H, b[:1], 0))
        else:
            th = rng.uniform(0, 6.28)
            out.append((np.diag([np.exp(-0.5j * th), np.exp(0.5j * th)]), b[:1], 0))
    return out


def test_sign_rules_and_their_switch(monkeypatch, plan_R):
    """Conditional sign flips are XORs of the sign bit ([zsign]); a sign controlled by a register bit with a pending
    exchange ([flip:zctl]) and a merged phase on the bit of one ([flip:phase] in two variants) commute over it.
    QBOT_B200_BRANCHY_SIGNS=1 restores FP64 negations under branches: same results, no XOR in the text."""
    import re
    seen = set()
    for seed in (2, 5, 8, 11):
        rng = np.random.default_rng(seed)
        n = int(rng.integers(12, 15))
        gl = _sign_heavy(rng, n, 70)
        psi = rand_ket(rng, n)
        ref = psi
        for m, tb, cm in gl:
            ref = oracle_apply_bits(ref, n, m, tb, cm)
        for branchy in (False, True):
            if branchy:
                monkeypatch.setenv('QBOT_B200_BRANCHY_SIGNS', '1')
            else:
                monkeypatch.delenv('QBOT_B200_BRANCHY_SIGNS', raising=False)
            text = ''.join(jit_emu.source_of(p) for f, _, p in jit_emu.plan(n, gl) if f)
            if branchy:
                assert 'QJ_XSIGN' not in text and '[flip:zctl]' not in text
            else:
                seen.update(re.findall(r'\[(zsign|flip:zctl|flip:phase)\]', text))
            out, _ = jit_emu.run(n, gl, psi)
            assert close(out, ref, 1e-12), (seed, n, branchy)
    assert {'zsign', 'flip:zctl', 'flip:phase'} <= seen, seen


def test_generated_source_matches_interpreter_semantics():
    """same plan, two executors: the op interpreter shared with the generic kernel
    (qb_tile_ops.h) and the generated straight-line code"""
    n = 15
    gl = plan_emu.circuit_to_bits(n, rc(n, 8, 3))
    psi = rand_ket(np.random.default_rng(5), n)
    a, _ = plan_emu.run(n, gl, psi)
    b, _ = jit_emu.run(n, gl, psi)
    assert close(b, a, 1e-14)


def test_structure_only_source():
    """angles are run-time coefficients: two circuits that differ only in their RZ angles and
    general 2x2 entries generate identical text"""
    n = 14
    a = plan_emu.circuit_to_bits(n, rc(n, 6, 9))
    b = []
    rng = np.random.default_rng(1)
    for m, tb, cm in a:
        m = np.asarray(m)
        if abs(m[0, 1]) == 0 and abs(m[0, 0] - 1) > 1e-9:      # an RZ: new angle
            th = rng.uniform(0.1, 6.0)
            m = np.diag([np.exp(-0.5j * th), np.exp(0.5j * th)])
        b.append((m, tb, cm))
    sa = [jit_emu.source_of(p) for f, _, p in jit_emu.plan(n, a) if f]
    sb = [jit_emu.source_of(p) for f, _, p in jit_emu.plan(n, b) if f]
    assert sa == sb and len(sa) > 0
    pa = [jit_emu.pool_of(p) for f, _, p in jit_emu.plan(n, a) if f]
    pb = [jit_emu.pool_of(p) for f, _, p in jit_emu.plan(n, b) if f]
    assert any(not np.array_equal(x, y) for x, y in zip(pa, pb))


def test_nvrtc_compiles_generated_kernels(tmp_path):
    """the real CUDA text through NVRTC for sm_100a, via the C ABI (qb_jit_check)"""
    import ctypes
    try:
        ctypes.CDLL('libnvrtc.so.12')
    except OSError:
        try:
            ctypes.CDLL('/usr/local/cuda/lib64/libnvrtc.so.12')
        except OSError:
            pytest.skip('libnvrtc not present')
    from qbot_b200 import _lib
    n = 24
    rng = np.random.default_rng(4)
    gl = plan_emu.circuit_to_bits(n, rc(n, 3, 24)) + tileable(random_gate_list(rng, n, 30))
    k = _lib.jit_check(n, gl, str(tmp_path))
    assert k >= 2
    cubins = sorted(tmp_path.glob('sweep_*.cubin'))
    assert len(cubins) == k and all(c.stat().st_size > 4096 for c in cubins)


def test_tile_bit_order_avoids_bank_class_exhaustion(monkeypatch):
    """The free tile bits of a sweep need not sit at ascending tile-local positions: the planner
    reorders them when a stage's register bits would exhaust a shared-memory bank class.  Find
    circuits where that happens, check they still execute correctly (generic op semantics and
    generated text), and that the benchmark plans have no conflicting stage at all."""
    import struct

    def bank_class(p):
        return 1 << p if p < 3 else 0 if p < 5 else 1 << ((p - 5) % 3)

    def stage_stats(program):
        nstages = struct.unpack_from('<IHH', program, 0)[1]
        hb = list(program[24:31])
        bad = 0
        for s in range(nstages):
            tpos = list(program[64 + 22 * s + 5:64 + 22 * s + 8])
            bad += sorted(bank_class(t) for t in tpos) != [1, 2, 4]
        return hb != sorted(hb), bad

    monkeypatch.setenv('QBOT_B200_PLAN_R', '5')
    found = 0
    for seed in range(40):
        n = 14
        gl = plan_emu.circuit_to_bits(n, rc(n, 8, 1000 + seed))
        steps = jit_emu.plan(n, gl)
        if not any(f and stage_stats(p)[0] for f, _, p in steps):
            continue
        found += 1
        psi = rand_ket(np.random.default_rng(seed), n)
        ref = psi
        for m, tb, cm in gl:
            ref = oracle_apply_bits(ref, n, m, tb, cm)
        out, _ = jit_emu.run(n, gl, psi)
        assert close(out, ref, 1e-12), seed
        if found >= 2:
            break
    assert found >= 1, "no circuit with a reordered sweep found -- widen the search"
    monkeypatch.setenv('QBOT_B200_PLAN_TRIALS', '32')
    for n, d, s in ((30, 20, 30), (34, 10, 34)):
        steps = jit_emu.plan(n, plan_emu.circuit_to_bits(n, rc(n, d, s)))
        assert sum(stage_stats(p)[1] for f, _, p in steps if f) == 0
