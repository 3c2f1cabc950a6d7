"""GPU: the fused tile engine (TMA-staged sweeps) against the oracle on mixed random circuits,
density matrices, branch batches; fused == unfused bit-for-bit is NOT required (different
operation order), 1e-12 relative is."""
import numpy as np
import pytest

from oracle import qbot_oracle as orc
from conftest import close
from test_planner import random_gate_list, oracle_apply_bits, rand_ket, rand_u

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def DS():
    from qbot_b200 import DeviceState
    return DeviceState


def test_fused_mixed_circuits(DS):
    rng = np.random.default_rng(11)
    for n in (12, 13, 15, 17):
        gl = random_gate_list(rng, n, 80)
        psi = rand_ket(rng, n)
        st = DS.from_host(psi)
        for m, tb, cm in gl:
            st.apply_gate_bits(m, tb, cm)
        got = np.asarray(st)
        stats = st.stats()
        assert stats['fused_passes'] > 0 and stats['fused_gates'] > 40, stats
        ref = psi
        for m, tb, cm in gl:
            ref = oracle_apply_bits(ref, n, m, tb, cm)
        assert close(got, ref, 1e-12), n


def test_fused_equals_unfused(DS):
    from qbot_b200.circuits import rc
    n = 18
    gates = rc(n, 30, 18)
    a, b = DS.zero_state(n), DS.zero_state(n)
    b.set_fusion(False)
    for g in gates:
        a.apply_gate(g.matrix(), g.target, g.controls)
        b.apply_gate(g.matrix(), g.target, g.controls)
    assert close(np.asarray(a), np.asarray(b), 1e-12)
    sa, sb = a.stats(), b.stats()
    assert sa['state_passes'] * 8 < sb['state_passes'], (sa, sb)


def test_fused_density_and_batch(DS):
    rng = np.random.default_rng(12)
    n = 7
    v, w = rand_ket(rng, n), rand_ket(rng, n)
    rho = 0.7 * np.outer(v, v.conj()) + 0.3 * np.outer(w, w.conj())
    st = DS.from_host(rho)
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
    assert st.stats()['fused_passes'] > 0
    # batch of 5 kets (not a power of two) with shared gates
    nb, B = 13, 5
    kets = np.stack([rand_ket(rng, nb) for _ in range(B)])
    bs = DS.from_kets(kets)
    gl = random_gate_list(rng, nb, 40)
    for m, tb, cm in gl:
        bs.apply_gate_bits(m, tb, cm)
    refs = []
    for kk in kets:
        r = kk
        for m, tb, cm in gl:
            r = oracle_apply_bits(r, nb, m, tb, cm)
        refs.append(r)
    assert close(np.asarray(bs), np.stack(refs), 1e-12)


def test_plan_cache_replay(DS):
    from qbot_b200.circuits import rc
    n = 14
    gates = rc(n, 8, 3)
    st = DS.zero_state(n)
    psi = np.zeros(1 << n, dtype=complex)
    psi[0] = 1
    for rep in range(3):
        for g in gates:
            st.apply_gate(g.matrix(), g.target, g.controls)
            psi = orc.ket_apply(psi, n, g.target, g.matrix(), g.controls)
        st.flush()
    assert close(np.asarray(st), psi, 1e-12)


def test_apply_circuit_equals_per_gate_calls(DS):
    """qb_apply_gates (one call for a whole gate list) == the same gates through qb_apply_gate"""
    from qbot_b200.circuits import rc
    n = 14
    gates = rc(n, 10, 77)
    items = [(g.matrix(), g.target, g.controls) for g in gates]
    a, b = DS.zero_state(n), DS.zero_state(n)
    a.apply_circuit(DS.pack_circuit(n, items))
    for m, t, c in items:
        b.apply_gate(m, t, c)
    assert np.array_equal(np.asarray(a), np.asarray(b))
    with pytest.raises(IndexError):
        DS.pack_circuit(n, [(np.eye(4), n - 1, [])])
