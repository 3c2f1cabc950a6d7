"""The oracle against outputs of the real reference (tests/golden/, see make_golden.py).
CPU only.  This is what pins the oracle; the CUDA path is then checked against the oracle."""
import numpy as np
import pytest

from oracle import qbot_oracle as orc
from qbot_b200.circuits import rc
from conftest import close

BASES = None


def bases():
    s = 2 ** (-1 / 2)
    comp = [np.array([1, 0], dtype=complex), np.array([0, 1], dtype=complex)]
    hada = [s * np.array([1, 1], dtype=complex), s * np.array([1, -1], dtype=complex)]
    bell = [s * np.array(v, dtype=complex) for v in ([1, 0, 0, 1], [0, 1, 1, 0], [1, 0, 0, -1], [0, 1, -1, 0])]
    return {k: [orc.ket_to_density(x) for x in v] for k, v in dict(comp=comp, hada=hada, bell=bell).items()}


def test_gate_cases(golden):
    cs = golden.cases('gate')
    assert len(cs) > 40
    for c in cs:
        g, rho, out = golden.arr(c['g']), golden.arr(c['rho']), golden.arr(c['out'])
        u = orc.controlled_unitary(c['n'], c['controls'], c['t'], g) if c['controls'] else orc.embed_gate(c['n'], c['t'], g)
        assert close(orc.conjugate(u, rho), out), c
        # tensor-form restatements agree with the matrix form
        assert close(orc.dm_apply(rho, c['n'], c['t'], g, c['controls']), out, 1e-11), c
        assert close(orc.reference_style_gate(rho, c['n'], c['t'], g, c['controls']), out), c


def test_ket_path_is_consistent_with_density(golden):
    rng = np.random.default_rng(5)
    for c in golden.cases('gate')[:60]:
        n = c['n']
        psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
        psi /= np.linalg.norm(psi)
        g = golden.arr(c['g'])
        u = orc.controlled_unitary(n, c['controls'], c['t'], g)
        assert close(orc.ket_apply(psi, n, c['t'], g, c['controls']), u @ psi, 1e-12)


def test_swap_and_shift(golden):
    for c in golden.cases('swap'):
        u = orc.swap_unitary(c['n'], c['a'], c['b'])
        assert close(orc.conjugate(u, golden.arr(c['rho'])), golden.arr(c['out'])), c
    for c in golden.cases('shift'):
        assert np.array_equal(orc.shift_unitary(c['n'], c['up'], c['shifts']).real, golden.arr(c['u'])), c


def test_ptrace(golden):
    for c in golden.cases('ptrace'):
        a, b = orc.ptrace_arbitrary(golden.arr(c['rho']), c['n'], c['qubits'])
        assert close(a, golden.arr(c['a'])) and close(b, golden.arr(c['b'])), c


def test_interweave_and_replace(golden):
    for c in golden.cases('interweave'):
        assert close(orc.interweave(golden.arr(c['a']), golden.arr(c['b']), c['pos']), golden.arr(c['out'])), c
    for c in golden.cases('replace'):
        assert close(orc.replace_arbitrary(golden.arr(c['rho']), golden.arr(c['new']), c['targets']), golden.arr(c['out'])), c


def test_measure(golden):
    bs = bases()
    cs = golden.cases('measure')
    assert len(cs) >= 30
    for c in cs:
        tg = c['targets_as_given']
        if tg is not None and c['targets_is_set']:
            tg = set(tg)
        r = orc.measure(golden.arr(c['rho']), bs[c['basis']], tg, c['return_state'])
        assert np.allclose(r['probs'], c['probs'], rtol=0, atol=2e-15), c     # rounded to 15 dp on both sides
        assert len(r['probs']) == len(c['probs'])
        assert close(r['unMeasuredDensity'], golden.arr(c['unmeasured'])), c
        if c['return_state']:
            assert close(r['newState'], golden.arr(c['new_state'])), c
        else:
            assert r['newState'] is None


def test_ensemble(golden):
    for c in golden.cases('ensemble'):
        out = orc.ensemble(c['probs'], [golden.arr(k) for k in c['rhos']])
        assert np.array_equal(out, golden.arr(c['out'])), c      # same accumulation order -> bit exact


def test_probval_rules(golden):
    for c in golden.probval:
        if c['kind'] == 'normalize':
            vals = [tuple(v) for v in c['values']] if c['tuple_values'] else c['values']
            p, v = orc.probval_normalize(c['probs'], vals)
            assert p == c['out_probs']
            assert [list(x) if isinstance(x, tuple) else x for x in v] == c['out_values']
        elif c['kind'] == 'fanout':
            order = orc.fan_out(c['lens'])
            a_vals, b_vals = [0, 1], [10, 20, 30]
            got = [[a_vals[i], 'k', b_vals[j]] for (i, j) in order]
            assert got == c['out_values']


def test_rc_circuits_match_reference(golden):
    for key in [k for k in golden.rc.files if not k.endswith('_info')]:
        _, n, depth, seed = key.split('_')
        n, depth, seed = int(n), int(depth), int(seed)
        gates = rc(n, depth, seed)
        assert len(gates) == int(golden.rc[key + '_info'][0])
        psi = np.zeros(1 << n, dtype=complex)
        psi[0] = 1
        for g in gates:
            psi = orc.ket_apply(psi, n, g.target, g.matrix(), g.controls)
        assert close(orc.ket_density(psi), golden.rc[key], 1e-12), key


def test_measure_collapse_is_product_state():
    # a Bell pair measured on qubit 0 collapses to I/4, not to the classical mixture
    bs = bases()
    r = orc.measure(bs['bell'][0], bs['comp'], [0], True)
    assert r['probs'] == [0.5, 0.5]
    assert close(r['newState'], np.eye(4) / 4)


def test_basis_weights_equals_the_literal_outcome_loop():
    """oracle.basis_weights (rotation form, used as the checker at 9-13 measured qubits) against the
    literal restatement of measurement.py:147-155 (abs(trace(rho_A @ P_i)) per outcome), all three
    DSL bases, full / partial / non-contiguous targets, density matrices and kets"""
    rng = np.random.default_rng(5)
    r2 = 2 ** -0.5
    bases = {'comp': [np.array([1, 0]), np.array([0, 1])], 'hada': [r2 * np.array([1, 1]), r2 * np.array([1, -1])],
             'bell': [r2 * np.array([1, 0, 0, 1]), r2 * np.array([0, 1, 1, 0]), r2 * np.array([1, 0, 0, -1]), r2 * np.array([0, 1, -1, 0])]}
    for n, t in ((1, [0]), (4, [0, 1, 2, 3]), (5, [1, 3]), (6, [0, 2, 3, 5]), (5, [0, 1, 2, 4]), (7, [1, 2, 4, 6])):
        d = 1 << n
        rho = np.zeros((d, d), dtype=complex)
        for _ in range(3):
            v = rng.normal(size=d) + 1j * rng.normal(size=d)
            v /= np.linalg.norm(v)
            rho += np.outer(v, v.conj()) / 3
        psi = rng.normal(size=d) + 1j * rng.normal(size=d)
        psi /= np.linalg.norm(psi)
        for name, kets in bases.items():
            b = orc.ilog2(len(kets[0]))
            if len(t) % b:
                continue
            dens = [np.outer(k, k).astype(complex) for k in kets]
            sys_a = rho if len(t) == n else orc.ptrace_arbitrary(rho, n, t)[0]
            f = len(t) // b
            lit = np.array([abs(np.trace(np.matmul(sys_a, orc.basis_projector(f, i, dens)[0]))) for i in range(len(dens) ** f)])
            assert np.max(np.abs(orc.basis_weights(rho, n, t, kets) - lit)) < 1e-14, (n, t, name)
            assert np.max(np.abs(orc.basis_weights(psi, n, t, kets) - orc.basis_weights(np.outer(psi, psi.conj()), n, t, kets))) < 1e-14


def test_inplace_ket_update_is_the_ket_update():
    """oracle.ket_apply_inplace (used by the GPU tests at 26 qubits) against oracle.ket_apply."""
    import numpy as np
    from oracle import qbot_oracle as orc
    from qbot_b200 import circuits
    n = 9
    rng = np.random.default_rng(5)
    psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    a, b = psi.copy(), psi.copy()
    for g in circuits.rc(n, 6, 9):
        a = orc.ket_apply(a, n, g.target, g.matrix(), g.controls)
        b = orc.ket_apply_inplace(b, n, g.target, g.matrix(), g.controls)
    u = np.linalg.qr(rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)))[0]
    a = orc.ket_apply(a, n, 3, u, [0])
    b = orc.ket_apply_inplace(b, n, 3, u, [0])
    assert np.max(np.abs(a - b)) <= 4e-16 * np.max(np.abs(a))
