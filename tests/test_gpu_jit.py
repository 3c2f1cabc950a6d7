"""GPU: structure-specialised (NVRTC) sweep kernels against the oracle and against the generic
paths -- forced specialisation (mode 2), the default repeat-triggered one, coefficient reuse of a
cached kernel, density matrices and branch batches."""
import numpy as np
import pytest

from oracle import qbot_oracle as orc
from conftest import close
from test_planner import random_gate_list, oracle_apply_bits, rand_ket, rand_u
from test_jit_codegen import tileable

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def DS():
    from qbot_b200 import DeviceState
    return DeviceState


@pytest.fixture(params=['5', '4'], ids=['R5', 'R4'])
def jit_R(request, monkeypatch):
    """plan shape of the specialised sweeps: 32 amplitudes per thread (default) or the generic 16"""
    monkeypatch.setenv('QBOT_B200_JIT_R', request.param)
    return int(request.param)


def test_jit_mixed_circuits_vs_oracle(DS, jit_R):
    rng = np.random.default_rng(21)
    for n in (12, 13, 15, 17):
        gl = random_gate_list(rng, n, 80)
        psi = rand_ket(rng, n)
        st = DS.from_host(psi)
        st.set_jit(2)
        for m, tb, cm in gl:
            st.apply_gate_bits(m, tb, cm)
        got = np.asarray(st)
        stats = st.stats()
        assert stats['jit_passes'] > 0 and stats['jit_passes'] == stats['fused_passes'], stats
        ref = psi
        for m, tb, cm in gl:
            ref = oracle_apply_bits(ref, n, m, tb, cm)
        assert close(got, ref, 1e-12), n


def test_jit_equals_unfused_rc(DS, jit_R):
    from qbot_b200.circuits import rc
    n = 22
    gates = rc(n, 12, 22)
    a, b = DS.zero_state(n), DS.zero_state(n)
    a.set_jit(2)
    b.set_fusion(False)
    for g in gates:
        a.apply_gate(g.matrix(), g.target, g.controls)
        b.apply_gate(g.matrix(), g.target, g.controls)
    assert close(np.asarray(a), np.asarray(b), 1e-12)
    assert a.stats()['jit_passes'] > 0


def test_jit_on_repeat_and_coefficient_reuse(DS):
    """default mode: the second flush of the same gate list compiles; a circuit with the same
    structure but other angles then hits the kernel cache"""
    from qbot_b200 import _lib
    from qbot_b200.circuits import rc
    n = 22
    gates = rc(n, 4, 5)
    st = DS.zero_state(n)
    psi = np.zeros(1 << n, dtype=complex)
    psi[0] = 1
    before = _lib.jit_info()
    for rep in range(3):
        for g in gates:
            st.apply_gate(g.matrix(), g.target, g.controls)
        st.flush()
    for rep in range(3):
        for g in gates:
            psi = orc.ket_apply(psi, n, g.target, g.matrix(), g.controls)
    assert close(np.asarray(st), psi, 1e-12)
    s = st.stats()
    assert 0 < s['jit_passes'] < s['fused_passes'], s            # first run interpreted, later ones specialised
    mid = _lib.jit_info()
    assert mid['kernels_compiled'] > before['kernels_compiled']
    # same structure, different RZ angles
    rng = np.random.default_rng(0)
    mats = []
    for g in gates:
        m = g.matrix()
        if abs(m[0, 1]) == 0 and abs(m[0, 0] - 1) > 1e-9:
            th = rng.uniform(0.1, 6.0)
            m = np.diag([np.exp(-0.5j * th), np.exp(0.5j * th)])
        mats.append(m)
    st2 = DS.zero_state(n)
    st2.set_jit(2)
    psi2 = np.zeros(1 << n, dtype=complex)
    psi2[0] = 1
    for g, m in zip(gates, mats):
        st2.apply_gate(m, g.target, g.controls)
        psi2 = orc.ket_apply(psi2, n, g.target, m, g.controls)
    assert close(np.asarray(st2), psi2, 1e-12)
    after = _lib.jit_info()
    assert after['kernels_compiled'] == mid['kernels_compiled'] and after['cache_hits'] > mid['cache_hits'], (mid, after)


def test_jit_density_and_batch(DS, jit_R):
    rng = np.random.default_rng(22)
    n = 7
    v, w = rand_ket(rng, n), rand_ket(rng, n)
    rho = 0.7 * np.outer(v, v.conj()) + 0.3 * np.outer(w, w.conj())
    st = DS.from_host(rho)
    st.set_jit(2)
    ref = rho
    for _ in range(25):
        k = int(rng.integers(1, 3))
        t = int(rng.integers(0, n - k + 1))
        free = [q for q in range(n) if q < t or q >= t + k]
        cs = [int(c) for c in rng.choice(free, size=int(rng.integers(0, 3)), replace=False)]
        g = rand_u(rng, k)
        st.apply_gate(g, t, cs)
        ref = orc.conjugate(orc.controlled_unitary(n, cs, t, g), ref)
    assert close(np.asarray(st), ref, 1e-12)
    assert st.stats()['jit_passes'] > 0
    nb, B = 13, 5
    kets = np.stack([rand_ket(rng, nb) for _ in range(B)])
    bs = DS.from_kets(kets)
    bs.set_jit(2)
    gl = tileable(random_gate_list(rng, nb, 40))
    for m, tb, cm in gl:
        bs.apply_gate_bits(m, tb, cm)
    refs = []
    for kk in kets:
        r = kk
        for m, tb, cm in gl:
            r = oracle_apply_bits(r, nb, m, tb, cm)
        refs.append(r)
    assert close(np.asarray(bs), np.stack(refs), 1e-12)
    assert bs.stats()['jit_passes'] > 0


def test_jit_large_coefficient_pool(DS, jit_R):
    """a sweep with more run-time coefficients than fit the kernel-parameter struct (> 480
    doubles: ~60 general 2x2 gates) reads them from a device array instead"""
    rng = np.random.default_rng(31)
    n = 13
    psi = rand_ket(rng, n)
    st = DS.from_host(psi)
    st.set_jit(2)
    gl = []
    for i in range(62):
        gl.append((rand_u(rng, 1), [int(rng.integers(0, 12))], 0))
    for m, tb, cm in gl:
        st.apply_gate_bits(m, tb, cm)
    got = np.asarray(st)
    assert st.stats()['jit_passes'] >= 1
    ref = psi
    for m, tb, cm in gl:
        ref = oracle_apply_bits(ref, n, m, tb, cm)
    assert close(got, ref, 1e-12)


def test_lazy_basis_state_initialisation():
    """A large single-branch ket initialised to a basis state (qb_init_basis, or a product of basis kets) is not
    written until something needs it: the first specialised sweep starts from the known amplitudes (qj_kernel's
    virtual-basis variant), every other consumer makes the library write the state first.  All routes against the
    oracle / the definition."""
    import numpy as np
    from oracle import qbot_oracle as orc
    from qbot_b200 import DeviceState, circuits
    n = 22
    # (1) read before any gate: probabilities, a range, a clone
    st = DeviceState.zero_state(n)
    assert st.probs([0, n - 1]).tolist() == [1.0, 0.0, 0.0, 0.0]
    st = DeviceState.zero_state(n)
    assert st.download_range(0, 2).tolist() == [1.0, 0.0]
    st = DeviceState.zero_state(n)
    c = st.clone()
    assert c.download_range(0, 1)[0] == 1.0 and float(c.norm2()[0]) == 1.0
    # (2) product of basis kets with some |1> factors, specialised sweeps at first sight
    ones = [1, 7, n - 1]
    factors = [np.array([0, 1], dtype=complex) if q in ones else np.array([1, 0], dtype=complex) for q in range(n)]
    index = sum(1 << (n - 1 - q) for q in ones)
    gates = circuits.rc(n, 3, 77)
    psi = np.zeros(1 << n, dtype=complex)
    psi[index] = 1
    for g in gates:
        orc.ket_apply_inplace(psi, n, g.target, g.matrix(), g.controls)
    for jit, fusion in ((2, True), (0, True), (2, False)):
        st = DeviceState.product(factors)
        st.set_jit(jit)
        st.set_fusion(fusion)
        for g in gates:
            st.apply_gate(g.matrix(), g.target, g.controls)
        got = np.asarray(st)
        assert np.max(np.abs(got - psi)) < 1e-12 * np.max(np.abs(psi)), (jit, fusion)
        if jit == 2 and fusion:
            s = st.stats()
            assert s['jit_passes'] == s['fused_passes'] > 0
    # (3) a one-gate kernel first (dense 3-qubit block: not fused), then sweeps
    u = np.linalg.qr(np.random.default_rng(3).normal(size=(8, 8)) + 1j * np.random.default_rng(4).normal(size=(8, 8)))[0]
    st = DeviceState.zero_state(n)
    st.set_jit(2)
    st.apply_gate(u, 4)
    ref = np.zeros(1 << n, dtype=complex)
    ref[0] = 1
    ref = orc.ket_apply(ref, n, 4, u, [])
    for g in gates[:20]:
        st.apply_gate(g.matrix(), g.target, g.controls)
        orc.ket_apply_inplace(ref, n, g.target, g.matrix(), g.controls)
    assert np.max(np.abs(np.asarray(st) - ref)) < 1e-12 * np.max(np.abs(ref))
    # (4) the same state object initialised again after use
    st = DeviceState.zero_state(n)
    st.set_jit(2)
    for rep in range(2):
        for g in gates:
            st.apply_gate(g.matrix(), g.target, g.controls)
        st.flush()
    from qbot_b200 import _lib
    _lib.call('qb_init_basis', st._h, index)
    st._dirty()
    for g in gates:
        st.apply_gate(g.matrix(), g.target, g.controls)
    assert np.max(np.abs(np.asarray(st) - psi)) < 1e-12 * np.max(np.abs(psi))
