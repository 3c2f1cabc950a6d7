"""GPU: row a8 -- outcome weights in any measurement basis, the collapsed system and the product-state
collapse on the device, against the oracle's restatement of
qbot/measurement.py:88-165.  No host matmul is involved on the product side: qb_probs_basis rotates a
scratch copy with the gate kernels and bins its diagonal, qb_init_diag + qb_apply_gate_rc build
sum_i p_i P_i.  Tolerance 1e-12 relative to the largest expected entry; outcome order exact."""
import numpy as np
import pytest

from oracle import qbot_oracle as orc
from conftest import close
from test_gpu_kernels import rand_dm, rand_ket

pytestmark = pytest.mark.gpu

R2 = 2 ** -0.5
KETS = {'comp': [np.array([1, 0], dtype=complex), np.array([0, 1], dtype=complex)],
        'hada': [R2 * np.array([1, 1], dtype=complex), R2 * np.array([1, -1], dtype=complex)],
        'bell': [R2 * np.array([1, 0, 0, 1], dtype=complex), R2 * np.array([0, 1, 1, 0], dtype=complex),
                 R2 * np.array([1, 0, 0, -1], dtype=complex), R2 * np.array([0, 1, -1, 0], dtype=complex)]}


@pytest.fixture(scope='module')
def DS():
    from qbot_b200 import DeviceState
    return DeviceState


def test_probs_basis_density_matrix_small_against_the_literal_loop(DS):
    rng = np.random.default_rng(41)
    for n, targets in ((1, [0]), (2, [0, 1]), (4, [0, 1, 2, 3]), (5, [1, 3]), (6, [0, 2, 3, 5]), (7, [1, 2, 4, 6])):
        rho = rand_dm(rng, n)
        st = DS.from_host(rho)
        for name, kets in KETS.items():
            b = orc.ilog2(kets[0].shape[0])
            if len(targets) % b:
                continue
            dens = [np.outer(k, k) for k in kets]
            sys_a = rho if len(targets) == n else orc.ptrace_arbitrary(rho, n, targets)[0]
            f = len(targets) // b
            want = np.array([abs(np.trace(np.matmul(sys_a, orc.basis_projector(f, i, dens)[0]))) for i in range(len(dens) ** f)])
            got = st.probs_basis(targets, kets)
            assert close(got, want), (n, targets, name)
        assert close(np.asarray(st), rho, 0.0)          # the register is untouched (peek semantics)


@pytest.mark.parametrize('basis', ['hada', 'bell'])
def test_probs_basis_9_to_12_targets(DS, basis):
    """the sizes the host loop could not do (2^m outcomes x 2^m x 2^m matmuls): 10 of 11 and all 12
    qubits of a density matrix, 12 of 20 qubits of a ket"""
    rng = np.random.default_rng(42)
    kets = KETS[basis]
    for n, targets in ((11, [0, 1, 2, 3, 4, 6, 7, 8, 9, 10]), (12, list(range(12)))):
        rho = rand_dm(rng, n, rank=2)
        st = DS.from_host(rho)
        got = st.probs_basis(targets, kets)
        want = orc.basis_weights(rho, n, targets, kets)
        assert got.shape == want.shape and close(got, want), (n, basis)
        del st
    n, targets = 20, [0, 2, 3, 5, 8, 9, 11, 12, 14, 16, 18, 19]
    psi = rand_ket(rng, n)
    st = DS.from_host(psi)
    got = st.probs_basis(targets, kets)
    assert close(got, orc.basis_weights(psi, n, targets, kets)), basis
    assert close(np.asarray(st), psi, 0.0)


def test_measure_op_all_bases_against_the_oracle(DS):
    """ops.measure end to end on the device: weights, outcome order, unmeasured density, collapse"""
    from qbot_b200.host.interp import Interpreter
    from qbot_b200.host.namespace import globalNameSpace as gns
    measure = Interpreter(DS).state_ops['measure']
    rng = np.random.default_rng(43)
    for n, targets in ((3, None), (4, [2, 0]), (5, [1, 3]), (6, {0, 2, 3, 5}), (6, [4, 5]), (8, [7, 1, 2, 4])):
        rho = rand_dm(rng, n)
        for name in ('comp', 'hada', 'bell'):
            basis = gns[name]
            m = n if targets is None else len(set(targets))
            if m % basis.numQubits:
                continue
            r = measure(DS.from_host(rho), basis, targets, True)
            want = orc.measure(rho, basis.density, targets, True, basis.ketSymbols)
            assert np.allclose(r.probs, want['probs'], rtol=0, atol=1e-12), (n, targets, name)
            assert list(r.basisSymbols) == want['basisSymbols']
            assert close(np.asarray(r.unMeasuredDensity), want['unMeasuredDensity'])
            assert close(np.asarray(r.newState), want['newState']), (n, targets, name)


def test_diagonal_and_row_column_gates(DS):
    rng = np.random.default_rng(44)
    for n in (1, 3, 6):
        w = rng.random(1 << n)
        st = DS.diagonal(w)
        assert np.array_equal(np.asarray(st), np.diag(w).astype(complex))
        r = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)) if n >= 2 else rng.normal(size=(2, 2)) + 0j
        c = rng.normal(size=r.shape) + 1j * rng.normal(size=r.shape)
        t = 0 if n < 3 else 1
        st.apply_gate_rc(r, c, t)
        big_r, big_c = orc.embed_gate(n, t, r), orc.embed_gate(n, t, c)
        assert close(np.asarray(st), big_r @ np.diag(w) @ big_c.T), n


def test_broadcast_then_per_branch_views(DS):
    """the fan-out copy must have landed before a single-branch view (own stream)
    applies its gate -- mixed gate sizes per branch take that path"""
    rng = np.random.default_rng(45)
    n = 12
    rho = rand_dm(rng, 6)
    base = DS.from_host(rho)
    for _ in range(5):
        batch = base.broadcast(3)
        g1, g2 = np.array([[0, 1], [1, 0]], dtype=complex), np.kron(KETS['hada'][0].reshape(2, 1) @ KETS['hada'][0].reshape(1, 2) * 2 - np.eye(2), np.eye(2))
        items = [(g1, [0], []), (g2, [2, 3], [5]), None]
        batch.apply_branch_gates(items)
        got = np.asarray(batch)
        assert close(got[0], orc.dm_apply(rho, 6, 0, g1, []))
        assert close(got[1], orc.dm_apply(rho, 6, 2, g2, [5]))
        assert close(got[2], rho, 0.0)


def test_probval_gate_on_a_ket_register_on_device(DS):
    import qbot_b200
    ket = "qset np_array([1, 0, 0, 0, 0, 0, 0, 0]) * (1+0j)\ngate hadamardGate ; 2\n"
    dm = "qset tensorExp(comp[0], 3)\ngate hadamardGate ; 2\n"
    for body in ("gate pauliXGate ; 0 ; [2] ; ProbVal([.5, .5], [True, False])\n",
                 "gate hadamardGate ; ProbVal([.25, .75], [0, 1])\n",
                 "gate ProbVal([.3, .7], [pauliXGate, hadamardGate]) ; 1 ; [2]\n",
                 "swap ProbVal([.5, .5], [0, 1]) ; 2\n"):
        a = qbot_b200.executeTxt(ket + body, state_cls=DS)
        b = qbot_b200.executeTxt(dm + body, state_cls=DS)
        ra, rb = np.asarray(a['state']), np.asarray(b['state'])
        assert ra.shape == (8, 8) and close(ra, rb), body
    # the C ABI itself refuses to sum kets
    from qbot_b200._lib import QbotB200Error
    k = DS.from_host(np.array([1, 0], dtype=complex))
    with pytest.raises(QbotB200Error):
        k.broadcast(2).mix_branches([0.5, 0.5])
