"""GPU: BASELINE.json configs 3 and 4 at FULL size, through size-independent properties plus
oracle parity on sampled pieces (config 2 at full size lives in test_gpu_kernels.py, config 5 in
test_gpu_sharded.py / bench.py).

Config 3: 12-qubit density matrix (2^24 entries), random circuit, mid-circuit measurement with
          the reference's product-state collapse, partial trace (`disc`).
Config 4: ProbVal batch of 4096 branch kets at 16 qubits, shared circuit + per-branch gate +
          measurement probabilities of 4 targets.
"""
import numpy as np
import pytest

from oracle import qbot_oracle as orc
from conftest import close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def DS():
    from qbot_b200 import DeviceState
    return DeviceState


def test_config3_density_matrix_12q(DS):
    from qbot_b200 import DM, KET
    from qbot_b200.circuits import rc
    n = 12
    gates = rc(n, 20, 12)
    rho = DS.zero_state(n, kind=DM)
    ket = DS.zero_state(n, kind=KET)
    psi = np.zeros(1 << n, dtype=complex)
    psi[0] = 1
    for g in gates:
        rho.apply_gate(g.matrix(), g.target, g.controls)
        ket.apply_gate(g.matrix(), g.target, g.controls)
        psi = orc.ket_apply(psi, n, g.target, g.matrix(), g.controls)
    # U rho U^dagger of a pure state == outer(psi, conj psi): DM path vs ket path vs oracle ket
    assert close(np.asarray(ket), psi, 1e-12)
    r = np.asarray(rho)
    assert r.shape == (1 << n, 1 << n)
    assert close(r, np.outer(psi, psi.conj()), 1e-12)
    assert abs(np.trace(r) - 1) < 1e-12
    # measurement probabilities of 2 and 4 targets == oracle on the ket
    for targets in ([2, 7], [0, 3, 8, 11]):
        assert close(rho.probs(targets).reshape(-1), orc.ket_probs(psi, n, targets), 1e-12)
    # partial trace (disc of 4 qubits): keep the other 8, compare with the einsum definition
    drop = [1, 4, 6, 10]
    keep = [q for q in range(n) if q not in drop]
    got = np.asarray(rho.ptrace_keep(keep))
    t = psi.reshape((2,) * n)
    ref = np.tensordot(t, t.conj(), axes=(drop, drop)).reshape(1 << len(keep), 1 << len(keep))
    assert close(got, ref, 1e-12)
    assert abs(np.trace(got) - 1) < 1e-12


def test_config3_script_with_measurement_branches():
    """12-qubit DSL program: gates, mid-circuit meas (reference collapse), ProbVal-target gate,
    disc; checked against the same program at full size through invariants and against the
    reference's semantics on the measured marginals."""
    import qbot_b200
    from qbot_b200.circuits import rc
    n = 12
    lines = ["qset tensorExp(comp[0], %d)" % n]
    lines += [g.dsl() for g in rc(n, 6, 3)]
    lines.append("peek before ; comp ; [3, 9]")
    lines.append("meas m1 ; comp ; [3, 9]")
    lines.append("gate hadamardGate ; ProbVal([.5,.5],[0,5])")
    lines.append("meas m2 ; comp ; [0]")
    lines.append("disc [1, 2, 4, 6]")
    ns = qbot_b200.executeTxt("\n".join(lines))
    p1 = np.array(ns['m1'].probs)
    assert abs(p1.sum() - 1) < 1e-12 and np.allclose(p1, np.array(ns['before'].probs), atol=1e-14)
    st = np.asarray(ns['state'])
    assert st.shape == (1 << 8, 1 << 8)
    assert abs(np.trace(st) - 1) < 1e-10
    assert np.allclose(st, st.conj().T, atol=1e-12)
    assert abs(sum(ns['m2'].probs) - 1) < 1e-12


def test_config4_branch_batch_4096x16(DS):
    from qbot_b200.circuits import rc
    n, B = 16, 4096
    rng = np.random.default_rng(16)
    th = rng.uniform(0, np.pi, size=(B, n))
    ph = rng.uniform(0, 2 * np.pi, size=(B, n))
    factors = np.stack([np.cos(th / 2), np.exp(1j * ph) * np.sin(th / 2)], axis=-1)     # [B, n, 2]
    st = DS.product_batch(factors)
    gates = rc(n, 10, 16)
    for g in gates:
        st.apply_gate(g.matrix(), g.target, g.controls)
    ang = rng.uniform(0, 2 * np.pi, size=B)
    tgt = rng.integers(0, n, size=B)
    mats = np.zeros((B, 2, 2), dtype=complex)
    mats[:, 0, 0] = np.exp(-0.5j * ang)
    mats[:, 1, 1] = np.exp(0.5j * ang)
    st.apply_gate_batched(mats, [int(t) for t in tgt])
    # a per-branch NON-diagonal gate too, on half of the branches (enable flags)
    X = np.array([[0, 1], [1, 0]], dtype=complex)
    en = (np.arange(B) % 2 == 0)
    st.apply_gate_batched(np.broadcast_to(X, (B, 2, 2)).copy(), [int((t + 3) % n) for t in tgt], enable=en)
    targets = [1, 6, 11, 15]
    probs = st.probs(targets)
    assert probs.shape == (B, 16)
    assert np.max(np.abs(probs.sum(axis=1) - 1)) < 1e-12
    assert np.max(np.abs(st.norm2() - 1)) < 1e-12
    # oracle parity on sampled branches
    for b in [0, 1, 777, 2048, 4095]:
        psi = np.array([1.0 + 0j])
        for q in range(n):
            psi = np.kron(psi, factors[b, q])
        for g in gates:
            psi = orc.ket_apply(psi, n, g.target, g.matrix(), g.controls)
        psi = orc.ket_apply(psi, n, int(tgt[b]), mats[b])
        if en[b]:
            psi = orc.ket_apply(psi, n, int((tgt[b] + 3) % n), X)
        assert close(probs[b], orc.ket_probs(psi, n, targets), 1e-12), b
        assert close(st.branch_view(b).to_host(), psi, 1e-12), b


def test_headline_30q_circuit_and_inverse(DS):
    """The benchmark workload at full size (30 qubits, 16 GiB): rc(30, 20) through the specialised
    sweeps, then the inverse circuit -- norm preserved, |0...0> recovered, marginals consistent;
    and the one-gate-per-pass kernels agree with the fused path on sampled amplitudes."""
    from qbot_b200.circuits import rc
    n = 30
    gates = rc(n, 20, 30)
    fwd = DS.pack_circuit(n, [(g.matrix(), g.target, g.controls) for g in gates])
    inv = DS.pack_circuit(n, [(g.matrix().conj().T, g.target, g.controls) for g in reversed(gates)])
    st = DS.zero_state(n)
    st.set_jit(2)
    st.apply_circuit(fwd)
    assert abs(st.norm2()[0] - 1.0) < 1e-11
    p = st.probs([0, 11, 29])
    assert abs(p.sum() - 1.0) < 1e-11
    idx = [0, 1, 12345, (1 << 29) + 7, (1 << 30) - 1] + [int(x) for x in np.random.default_rng(30).integers(0, 1 << 30, 59)]
    sample = np.array([st.download_range(i, 1)[0] for i in idx])
    assert st.stats()['jit_passes'] == st.stats()['fused_passes'] > 0
    # reference for the samples: the same circuit, one gate per pass (independent kernels)
    ref = DS.zero_state(n)
    ref.set_fusion(False)
    ref.apply_circuit(fwd)
    want = np.array([ref.download_range(i, 1)[0] for i in idx])
    ref_p = ref.probs([0, 11, 29])
    del ref
    # 1e-12 RELATIVE to the largest amplitude of the sample (|psi_i| ~ 3e-5 here), as BASELINE's north_star states it
    assert np.max(np.abs(sample - want)) < 1e-12 * np.max(np.abs(want))
    assert np.max(np.abs(p - ref_p)) < 1e-12 * np.max(ref_p)
    st.apply_circuit(inv)
    amp = st.download_range(0, 4)
    assert abs(amp[0] - 1.0) < 1e-10 and np.max(np.abs(amp[1:])) < 1e-10
    assert abs(st.probs([5])[0] - 1.0) < 1e-10


def test_specialised_sweeps_26q_against_the_oracle(DS):
    """The headline path (planner + NVRTC-specialised qj_kernel sweeps) against the ORACLE at 26 qubits
    (1 GiB ket on the host, oracle.ket_apply_inplace): every amplitude, 1e-12 relative to the largest."""
    from qbot_b200.circuits import rc
    n = 26
    gates = rc(n, 4, 26)
    st = DS.zero_state(n)
    st.set_jit(2)
    st.apply_circuit(DS.pack_circuit(n, [(g.matrix(), g.target, g.controls) for g in gates]))
    got = np.asarray(st)
    stats = st.stats()
    assert stats['jit_passes'] == stats['fused_passes'] > 0
    del st
    psi = np.zeros(1 << n, dtype=complex)
    psi[0] = 1
    for g in gates:
        orc.ket_apply_inplace(psi, n, g.target, g.matrix(), g.controls)
    scale = float(np.max(np.abs(psi)))
    got -= psi
    assert float(np.max(np.abs(got))) < 1e-12 * scale
