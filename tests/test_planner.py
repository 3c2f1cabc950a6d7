"""CPU: the fusion planner (qbot_b200/csrc/qb_plan.cpp) checked by executing its sweep programs
with the CPU plan emulator (same addressing and op semantics as the tile kernel) against the
oracle.  Covers gate reordering legality, stage / register-bit assignment, predicates on
register / thread / out-of-tile bits, phase merging and the unfused fallbacks."""
import numpy as np
import pytest

import plan_emu
from oracle import qbot_oracle as orc
from qbot_b200.circuits import rc
from conftest import close


def rand_ket(rng, n):
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    return v / np.linalg.norm(v)


def rand_u(rng, k):
    a = rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k))
    q, r = np.linalg.qr(a)
    return q * (np.diag(r) / np.abs(np.diag(r)))


def oracle_apply_bits(psi, n, m, tb, cmask):
    """Gate on arbitrary index bits through the oracle (reference qubit = n-1-bit)."""
    k = len(tb)
    qs = [n - 1 - b for b in tb]
    controls = [n - 1 - b for b in range(n) if (cmask >> b) & 1]
    if qs == list(range(qs[0], qs[0] + k)):
        return orc.ket_apply(psi, n, qs[0], m, controls)
    # non-contiguous / unordered targets: permute axes
    t = np.array(psi, dtype=complex).reshape((2,) * n)
    sl = [slice(None)] * n
    for c in controls:
        sl[c] = 1
    sub = t[tuple(sl)]
    ctl = sorted(controls)
    axes = [q - sum(1 for c in ctl if c < q) for q in qs]
    g = np.asarray(m, dtype=complex).reshape((2,) * (2 * k))
    res = np.tensordot(g, sub, axes=(list(range(k, 2 * k)), axes))
    res = np.moveaxis(res, list(range(k)), axes)
    t[tuple(sl)] = res
    return t.reshape(-1)


def random_gate_list(rng, n, count):
    X = np.array([[0, 1], [1, 0]], dtype=complex)
    H = np.array([[1, 1], [1, -1]], dtype=complex) * 2 ** -0.5
    out = []
    for _ in range(count):
        kind = int(rng.integers(0, 9))
        bits = [int(b) for b in rng.permutation(n)]
        nc = int(rng.integers(0, 3))
        if kind == 0:
            m, tb = H, bits[:1]
        elif kind == 1:
            m, tb = X, bits[:1]
        elif kind == 2:
            m, tb = np.diag(np.exp(1j * rng.uniform(0, 6.28, 2))), bits[:1]
        elif kind == 3:
            m, tb = rand_u(rng, 1), bits[:1]
        elif kind == 4:
            m, tb = rand_u(rng, 2), bits[:2]
        elif kind == 5:
            m, tb = np.diag(np.exp(1j * rng.uniform(0, 6.28, 4))), bits[:2]
        elif kind == 6:
            m, tb = np.diag(np.exp(1j * rng.uniform(0, 6.28, 8))), bits[:3]
        elif kind == 7:
            p = rng.permutation(4)
            m = np.zeros((4, 4), dtype=complex)
            m[np.arange(4), p] = np.exp(1j * rng.uniform(0, 6.28, 4))
            tb = bits[:2]
        else:
            m, tb, nc = rand_u(rng, 3), bits[:3], 0          # not tileable: unfused step
        cm = 0
        for c in bits[len(tb):len(tb) + nc]:
            cm |= 1 << c
        if kind == 2 and nc == 0 and rng.random() < 0.5:
            cm = 0
        out.append((m, tb, cm))
    return out


@pytest.mark.parametrize('M', [11, 12])
@pytest.mark.parametrize('merge', [True, False])
def test_random_mixed_circuits(M, merge):
    rng = np.random.default_rng(100 + M)
    for n in (12, 13, 15):
        gl = random_gate_list(rng, n, 60)
        psi = rand_ket(rng, n)
        out, st = plan_emu.run(n, gl, psi, M=M, merge=merge)
        ref = psi
        for m, tb, cm in gl:
            ref = oracle_apply_bits(ref, n, m, tb, cm)
        assert close(out, ref, 1e-12), (n, st)
        assert st['fused_gates'] + st['unfused_steps'] == len(gl)


@pytest.mark.parametrize('M', [11, 12])
def test_rc_circuits(M):
    for n, depth, seed in ((12, 10, 1), (14, 12, 2), (16, 6, 3)):
        gates = rc(n, depth, seed)
        rng = np.random.default_rng(seed)
        psi = rand_ket(rng, n)
        out, st = plan_emu.run(n, plan_emu.circuit_to_bits(n, gates), psi, M=M)
        ref = psi
        for g in gates:
            ref = orc.ket_apply(ref, n, g.target, g.matrix(), g.controls)
        assert close(out, ref, 1e-12), (n, st)
        assert st['unfused_steps'] == 0


def test_density_matrix_gate_pairs():
    """DM mode queues U on the row bits and conj(U) on the column bits."""
    rng = np.random.default_rng(7)
    n = 6
    dim = 1 << n
    v = rand_ket(rng, n)
    w = rand_ket(rng, n)
    rho = 0.6 * np.outer(v, v.conj()) + 0.4 * np.outer(w, w.conj())
    gl, ref = [], rho
    for _ in range(20):
        k = int(rng.integers(1, 3))
        t = int(rng.integers(0, n - k + 1))
        free = [q for q in range(n) if q < t or q >= t + k]
        cs = [int(c) for c in rng.choice(free, size=int(rng.integers(0, 3)), replace=False)]
        g = rand_u(rng, k)
        tb = [n - 1 - (t + j) for j in range(k)]
        cm = sum(1 << (n - 1 - c) for c in cs)
        gl.append((g, [b + n for b in tb], cm << n))
        gl.append((g.conj(), tb, cm))
        ref = orc.conjugate(orc.controlled_unitary(n, cs, t, g), ref)
    out, st = plan_emu.run(2 * n, gl, rho.reshape(-1))
    assert close(out.reshape(dim, dim), ref, 1e-12)


def test_headline_plans_are_compact():
    """Planning only (no execution) at the benchmark sizes: the whole circuit must fuse and the
    sweep count must stay far below the gate count."""
    for n, depth, seed, max_sweeps in ((30, 20, 30, 20), (20, 200, 20, 90), (34, 10, 34, 14)):
        gates = rc(n, depth, seed)
        _, st = plan_emu.run(n, plan_emu.circuit_to_bits(n, gates), None, execute=False)
        assert st['unfused_steps'] == 0 and st['fused_gates'] == len(gates)
        assert st['fused_sweeps'] <= max_sweeps, st
        assert st['max_program_bytes'] <= 8192


def test_small_states_are_not_tiled():
    gl = [(np.array([[0, 1], [1, 0]], dtype=complex), [3], 0)]
    _, st = plan_emu.run(8, gl, np.ones(256, dtype=complex), M=12)
    assert st['fused_sweeps'] == 0 and st['unfused_steps'] == 1


def test_plan_search_never_worse_and_still_correct(monkeypatch):
    """randomised plan search (qt_plan search_trials): trial 0 is the greedy plan, so the result
    can only have fewer steps; the searched plan executes to the same state"""
    n = 16
    gates = rc(n, 14, 5)
    gl = plan_emu.circuit_to_bits(n, gates)
    psi = rand_ket(np.random.default_rng(2), n)
    monkeypatch.setenv('QBOT_B200_PLAN_TRIALS', '1')
    base, st1 = plan_emu.run(n, gl, psi)
    monkeypatch.setenv('QBOT_B200_PLAN_TRIALS', '24')
    out, st2 = plan_emu.run(n, gl, psi)
    assert st2['steps'] <= st1['steps']
    assert st2['fused_gates'] == len(gl) and st2['unfused_steps'] == 0
    assert close(out, base, 1e-13)
    for nn, d, s in ((30, 20, 30), (34, 10, 34)):
        gl = plan_emu.circuit_to_bits(nn, rc(nn, d, s))
        monkeypatch.setenv('QBOT_B200_PLAN_TRIALS', '1')
        _, a = plan_emu.run(nn, gl, None, execute=False)
        monkeypatch.setenv('QBOT_B200_PLAN_TRIALS', '32')          # the engine's setting for states of >= 27 index bits
        _, b = plan_emu.run(nn, gl, None, execute=False)
        assert b['fused_sweeps'] <= a['fused_sweeps'], (a, b)
        assert b['fused_sweeps'] <= (12 if nn == 30 else 8)


def test_deep_plan_search_is_correct_and_not_worse(monkeypatch):
    """the deepest search level (QBOT_B200_PLAN_TRIALS >= 128: beam 24 x 10 plus the level below it), the engine's
    setting for specialised plans on states of >= 29 index bits: executes to the same state as the greedy plan,
    through the op interpreter and through the generated kernels, and never needs more sweeps than level 32"""
    import jit_emu
    for n, depth, seed in ((15, 16, 7), (14, 24, 8)):
        gates = rc(n, depth, seed)
        gl = plan_emu.circuit_to_bits(n, gates)
        psi = rand_ket(np.random.default_rng(seed), n)
        monkeypatch.setenv('QBOT_B200_PLAN_TRIALS', '1')
        base, st1 = plan_emu.run(n, gl, psi)
        monkeypatch.setenv('QBOT_B200_PLAN_TRIALS', '128')
        out, st2 = plan_emu.run(n, gl, psi)
        assert st2['steps'] <= st1['steps']
        assert close(out, base, 1e-13)
        monkeypatch.setenv('QBOT_B200_PLAN_R', '5')
        out5, _ = jit_emu.run(n, gl, psi)
        monkeypatch.delenv('QBOT_B200_PLAN_R')
        assert close(out5, base, 1e-12)
    for nn, d, s, most in ((30, 20, 30, 11), (31, 20, 7, 11), (34, 10, 34, 8)):
        gl = plan_emu.circuit_to_bits(nn, rc(nn, d, s))
        monkeypatch.setenv('QBOT_B200_PLAN_TRIALS', '32')
        _, a = plan_emu.run(nn, gl, None, execute=False)
        monkeypatch.setenv('QBOT_B200_PLAN_TRIALS', '128')
        _, b = plan_emu.run(nn, gl, None, execute=False)
        assert (b['fused_sweeps'], b.get('stages', 0)) <= (a['fused_sweeps'], a.get('stages', 0)), (a, b)
        assert b['fused_sweeps'] <= most, b


def test_peephole_rewrite_preserves_the_state(monkeypatch):
    """X (any controls) followed by H on its target is rewritten to H followed by a controlled Z
    (H X = Z H): same state, the diagonal gate needs no tile bit and costs a sign flip"""
    rng = np.random.default_rng(9)
    n = 13
    X = np.array([[0, 1], [1, 0]], dtype=complex)
    H = np.array([[1, 1], [1, -1]], dtype=complex) * 2 ** -0.5
    gl = []
    for _ in range(60):
        bits = [int(b) for b in rng.permutation(n)]
        kind = int(rng.integers(0, 5))
        if kind == 0:
            gl.append((X, bits[:1], sum(1 << c for c in bits[1:1 + int(rng.integers(0, 3))])))
            if rng.random() < 0.7:          # something harmless in between, then the Hadamard on the target
                gl.append((np.diag(np.exp(1j * rng.uniform(0, 6, 2))), [bits[4]], 0))
                gl.append((H, bits[:1], 0))
        elif kind == 1:
            gl.append((H, bits[:1], 0))
        elif kind == 2:
            gl.append((np.diag(np.exp(1j * rng.uniform(0, 6, 2))), bits[:1], 0))
        elif kind == 3:
            gl.append((X, bits[:1], 1 << bits[1]))
            gl.append((H, [bits[1]], 0))      # Hadamard on the CONTROL: blocks the rewrite across it
            gl.append((H, bits[:1], 0))
        else:
            gl.append((rand_u(rng, 1), bits[:1], 1 << bits[2]))
    psi = rand_ket(rng, n)
    ref = psi
    for m, tb, cm in gl:
        ref = oracle_apply_bits(ref, n, m, tb, cm)
    monkeypatch.delenv('QBOT_B200_NO_PEEPHOLE', raising=False)
    out, st = plan_emu.run(n, gl, psi)
    assert close(out, ref, 1e-12)
    monkeypatch.setenv('QBOT_B200_NO_PEEPHOLE', '1')
    out2, st2 = plan_emu.run(n, gl, psi)
    assert close(out2, ref, 1e-12)
    assert st['fused_gates'] == st2['fused_gates'] == len(gl)


def test_phase_ops_are_placed_on_a_minimum_set_of_stage_tails(monkeypatch):
    """Merged phase ops: the stage tails that carry one are a minimum set meeting every entry's commutation
    window (qb_plan.cpp place_diagonals); QBOT_B200_PHASE_GREEDY=1 restores the entry-by-entry placement.
    Both execute correctly and the minimum set never needs more ops."""
    from qbot_b200.circuits import rc
    fewer = 0
    for n, depth, seed in ((14, 10, 1), (15, 12, 2), (13, 20, 3), (16, 8, 4)):
        gl = plan_emu.circuit_to_bits(n, rc(n, depth, seed))
        psi = rand_ket(np.random.default_rng(seed), n)
        ref = psi
        for m, tb, cm in gl:
            ref = oracle_apply_bits(ref, n, m, tb, cm)
        ops = {}
        for greedy in (False, True):
            if greedy:
                monkeypatch.setenv('QBOT_B200_PHASE_GREEDY', '1')
            else:
                monkeypatch.delenv('QBOT_B200_PHASE_GREEDY', raising=False)
            out, st = plan_emu.run(n, gl, psi)
            assert close(out, ref, 1e-12), (n, greedy)
            ops[greedy] = (st['fused_sweeps'], st['stages'], st['ops'])
        if ops[False][:2] == ops[True][:2]:            # same sweeps and stages: only the phase ops differ
            assert ops[False][2] <= ops[True][2], ops
            fewer += ops[False][2] < ops[True][2]
    assert fewer >= 1
