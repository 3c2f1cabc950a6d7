"""GPU parity: every kernel behind the C ABI against the CPU oracle on seeded inputs, and
against the fixtures produced by the real reference (tests/golden).  Tolerance: 1e-12 relative
to the largest expected entry (BASELINE.json north_star: fp64, 1e-12 relative); index / order
results exact."""
import numpy as np
import pytest

from oracle import qbot_oracle as orc
from conftest import close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def DS():
    from qbot_b200 import DeviceState
    return DeviceState


def rand_ket(rng, n):
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    return v / np.linalg.norm(v)


def rand_dm(rng, n, rank=3):
    dim = 1 << n
    rho = np.zeros((dim, dim), dtype=complex)
    w = rng.random(rank) + 0.1
    w /= w.sum()
    for p in w:
        v = rand_ket(rng, n)
        rho += p * np.outer(v, v.conj())
    return rho


def rand_u(rng, k):
    a = rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k))
    q, r = np.linalg.qr(a)
    return q * (np.diag(r) / np.abs(np.diag(r)))


@pytest.mark.parametrize('fusion', [False, True])
def test_ket_dense_gates_all_positions(DS, fusion):
    rng = np.random.default_rng(1)
    for n in (1, 2, 3, 5, 8, 11):
        for k in (1, 2, 3, 4, 5):
            if k > n:
                continue
            for t in sorted(set([0, (n - k) // 2, n - k])):
                psi = rand_ket(rng, n)
                g = rand_u(rng, k)
                free = [q for q in range(n) if q < t or q >= t + k]
                nc = int(rng.integers(0, min(3, len(free)) + 1))
                controls = [int(c) for c in rng.choice(free, size=nc, replace=False)] if nc else []
                st = DS.from_host(psi)
                st.set_fusion(fusion)
                st.apply_gate(g, t, controls)
                assert close(np.asarray(st), orc.ket_apply(psi, n, t, g, controls)), (n, k, t, controls)


@pytest.mark.parametrize('fusion', [False, True])
def test_ket_structured_gates(DS, fusion):
    rng = np.random.default_rng(2)
    X = np.array([[0, 1], [1, 0]], dtype=complex)
    for n in (3, 6, 10, 13):
        psi = rand_ket(rng, n)
        st = DS.from_host(psi)
        st.set_fusion(fusion)
        ref = psi
        for _ in range(12):
            kind = rng.integers(0, 4)
            t = int(rng.integers(0, n))
            others = [q for q in range(n) if q != t]
            if kind == 0:      # diagonal with controls
                g = np.diag(np.exp(1j * rng.uniform(0, 6.28, 2)))
                cs = [int(c) for c in rng.choice(others, size=min(2, len(others)), replace=False)]
            elif kind == 1:    # X / CNOT / Toffoli (permutation)
                g, cs = X, [int(c) for c in rng.choice(others, size=int(rng.integers(0, 3)), replace=False)]
            elif kind == 2:    # 2-qubit diagonal
                t = int(rng.integers(0, n - 1))
                g, cs = np.diag(np.exp(1j * rng.uniform(0, 6.28, 4))), []
            else:              # monomial (phase * permutation) on 2 qubits
                t = int(rng.integers(0, n - 1))
                perm = rng.permutation(4)
                g = np.zeros((4, 4), dtype=complex)
                g[np.arange(4), perm] = np.exp(1j * rng.uniform(0, 6.28, 4))
                cs = []
            st.apply_gate(g, t, cs)
            ref = orc.ket_apply(ref, n, t, g, cs)
        assert close(np.asarray(st), ref), n


def test_large_gates_fallback(DS):
    rng = np.random.default_rng(3)
    for n, k, t in ((7, 6, 0), (8, 7, 1), (9, 6, 2)):
        psi = rand_ket(rng, n)
        g = rand_u(rng, k)
        st = DS.from_host(psi)
        st.apply_gate(g, t, [n - 1] if t + k < n else [])
        assert close(np.asarray(st), orc.ket_apply(psi, n, t, g, [n - 1] if t + k < n else [])), (n, k)


@pytest.mark.parametrize('fusion', [False, True])
def test_dense_blocks_on_the_tensor_cores(DS, fusion):
    """Row f3: dense 6..8-qubit blocks (qftGate(k), user unitaries; qbot/qgates.py:63-74, 161-182) run as a
    complex GEMM on the FP64 tensor cores (k_dense_mma) once the register has 12 free index bits."""
    from qbot_b200.host import hostmath as hm
    rng = np.random.default_rng(33)
    used = 0
    for n, k in ((12, 6), (13, 6), (14, 7), (15, 8), (16, 6), (16, 8)):
        for t in sorted(set([0, (n - k) // 2, n - k])):
            for g in (rand_u(rng, k), hm.qft(k)):
                psi = rand_ket(rng, n)
                free = [q for q in range(n) if q < t or q >= t + k]
                nc = int(rng.integers(0, 3))
                controls = [int(c) for c in rng.choice(free, size=nc, replace=False)] if nc and n - k - nc >= 12 - k else []
                st = DS.from_host(psi)
                st.set_fusion(fusion)
                st.reset_stats()
                st.apply_gate(g, t, controls)
                got = np.asarray(st)
                assert close(got, orc.ket_apply(psi, n, t, g, controls)), (n, k, t, controls)
                used += 1
    assert used
    # density matrix: U on the row bits, conj(U) on the column bits of vec(rho)
    n, k, t = 8, 6, 1
    rho = rand_dm(rng, n)
    g = rand_u(rng, k)
    sd = DS.from_host(rho)
    sd.set_fusion(fusion)
    sd.apply_gate(g, t, [])
    assert close(np.asarray(sd), orc.dm_apply(rho, n, t, g, []))
    # branch batch: the same block on every branch ket
    nb, n, k, t = 3, 13, 7, 2
    kets = np.stack([rand_ket(rng, n) for _ in range(nb)])
    sb = DS.from_kets(kets)
    sb.set_fusion(fusion)
    g = rand_u(rng, k)
    sb.apply_gate(g, t, [0])
    out = np.asarray(sb).reshape(nb, -1)
    for b in range(nb):
        assert close(out[b], orc.ket_apply(kets[b], n, t, g, [0])), b


def test_swap(DS):
    rng = np.random.default_rng(4)
    for n in (2, 5, 9):
        for _ in range(4):
            a, b = (int(x) for x in rng.integers(0, n, 2))
            psi = rand_ket(rng, n)
            st = DS.from_host(psi)
            st.apply_swap(a, b)
            assert np.array_equal(np.asarray(st), orc.ket_swap(psi, n, a, b)), (n, a, b)
            rho = rand_dm(rng, min(n, 5))
            m = min(n, 5)
            a, b = a % m, b % m
            sd = DS.from_host(rho)
            sd.apply_swap(a, b)
            assert np.array_equal(np.asarray(sd), orc.conjugate(orc.swap_unitary(m, a, b), rho))


@pytest.mark.parametrize('fusion', [False, True])
def test_density_matrix_conjugation(DS, fusion):
    rng = np.random.default_rng(5)
    for n in (1, 2, 4, 6):
        for k in (1, 2, 3):
            if k > n:
                continue
            t = int(rng.integers(0, n - k + 1))
            free = [q for q in range(n) if q < t or q >= t + k]
            cs = [int(c) for c in rng.choice(free, size=min(len(free), int(rng.integers(0, 3))), replace=False)]
            rho, g = rand_dm(rng, n), rand_u(rng, k)
            st = DS.from_host(rho)
            st.set_fusion(fusion)
            st.apply_gate(g, t, cs)
            u = orc.controlled_unitary(n, cs, t, g)
            assert close(np.asarray(st), orc.conjugate(u, rho)), (n, k, t, cs)


def test_golden_gate_and_swap_cases(DS, golden):
    for c in golden.cases('gate'):
        st = DS.from_host(golden.arr(c['rho']))
        st.apply_gate(golden.arr(c['g']), c['t'], c['controls'])
        assert close(np.asarray(st), golden.arr(c['out'])), c
    for c in golden.cases('swap'):
        st = DS.from_host(golden.arr(c['rho']))
        st.apply_swap(c['a'], c['b'])
        assert close(np.asarray(st), golden.arr(c['out'])), c


def test_probs_and_norm(DS):
    rng = np.random.default_rng(6)
    for n in (1, 4, 9, 12, 14):
        psi = rand_ket(rng, n) * 1.3
        st = DS.from_host(psi)
        assert abs(st.norm2()[0] - 1.69) < 1e-12
        for m in (1, 2, min(n, 5), n if n <= 12 else 3):
            if m > n:
                continue
            qs = [int(q) for q in rng.choice(n, size=m, replace=False)]
            assert close(st.probs(qs), orc.ket_probs(psi, n, qs)), (n, qs)
    for n in (1, 3, 6):
        rho = rand_dm(rng, n)
        sd = DS.from_host(rho)
        for m in (1, n):
            qs = [int(q) for q in rng.choice(n, size=m, replace=False)]
            a, _ = orc.ptrace_arbitrary(rho, n, qs) if m < n else (rho, None)
            # ptrace_arbitrary sorts: build the expected weights in listed order
            d = np.abs(np.diag(rho)).reshape((2,) * n)
            others = tuple(q for q in range(n) if q not in qs)
            d = d.sum(axis=others) if others else d
            asc = sorted(qs)
            exp = np.transpose(d, [asc.index(q) for q in qs]).reshape(-1)
            assert close(sd.probs(qs), exp, 1e-12), (n, qs)
        assert abs(sd.norm2()[0] - 1.0) < 1e-12


def test_ptrace_scatter_mix(DS, golden):
    for c in golden.cases('ptrace'):
        st = DS.from_host(golden.arr(c['rho']))
        keep = sorted(set(c['qubits']))
        rest = [q for q in range(c['n']) if q not in keep]
        assert close(np.asarray(st.ptrace_keep(keep)), golden.arr(c['a'])), c
        assert close(np.asarray(st.ptrace_keep(rest)), golden.arr(c['b'])), c
    for c in golden.cases('interweave'):
        a, b = DS.from_host(golden.arr(c['a'])), DS.from_host(golden.arr(c['b']))
        n = a.nq + b.nq
        pos = sorted(c['pos'])
        rest = [q for q in range(n) if q not in pos]
        assert close(np.asarray(DS.scatter_product(a, b, pos, rest)), golden.arr(c['out'])), c
    for c in golden.cases('ensemble'):
        sts = [DS.from_host(golden.arr(k)) for k in c['rhos']]
        assert np.array_equal(np.asarray(DS.mix(c['probs'], sts)), golden.arr(c['out'])), c   # bit exact
    rng = np.random.default_rng(7)
    rho = rand_dm(rng, 9)       # large traced space -> block-reduction kernel
    st = DS.from_host(rho)
    a, b = orc.ptrace_arbitrary(rho, 9, [2, 7])
    assert close(np.asarray(st.ptrace_keep([2, 7])), a) and close(np.asarray(st.ptrace_keep([0, 1, 3, 4, 5, 6, 8])), b)


def test_outer_and_product_init(DS):
    rng = np.random.default_rng(8)
    psi = rand_ket(rng, 5)
    st = DS.from_host(psi)
    assert close(np.asarray(st.outer(True)), np.outer(psi, psi.conj()))
    assert close(np.asarray(st.outer(False)), np.outer(psi, psi))
    vecs = [rand_ket(rng, 1) for _ in range(6)]
    exp = vecs[0]
    for v in vecs[1:]:
        exp = np.kron(exp, v)
    assert close(np.asarray(DS.product(vecs)), exp)
    from qbot_b200 import DM
    mats = [rand_dm(rng, 1) for _ in range(4)]
    exp = mats[0]
    for m in mats[1:]:
        exp = np.kron(exp, m)
    assert close(np.asarray(DS.product(mats, kind=DM)), exp)
    z = DS.zero_state(7)
    e = np.zeros(128, dtype=complex)
    e[0] = 1
    assert np.array_equal(np.asarray(z), e)


def test_batched_branch_gates(DS):
    rng = np.random.default_rng(9)
    n, B = 6, 37
    kets = np.stack([rand_ket(rng, n) for _ in range(B)])
    st = DS.from_kets(kets)
    # shared gate on every branch
    g = rand_u(rng, 2)
    st.apply_gate(g, 3, [0])
    ref = np.stack([orc.ket_apply(k, n, 3, g, [0]) for k in kets])
    assert close(np.asarray(st), ref)
    # per-branch gates, targets, controls, enable flags
    mats = np.stack([rand_u(rng, 1) for _ in range(B)])
    targets = [int(t) for t in rng.integers(0, n, B)]
    controls = [[int(c) for c in rng.choice([q for q in range(n) if q != t], size=int(rng.integers(0, 3)), replace=False)] for t in targets]
    enable = [bool(x) for x in rng.integers(0, 2, B)]
    st.apply_gate_batched(mats, targets, controls, enable)
    ref = np.stack([orc.ket_apply(ref[b], n, targets[b], mats[b], controls[b]) if enable[b] else ref[b] for b in range(B)])
    assert close(np.asarray(st), ref)
    p = st.probs([4, 1])
    assert close(p, np.stack([orc.ket_probs(r, n, [4, 1]) for r in ref]))
    # density batch + weighted reduction == the reference's per-branch loop + ensemble
    rho = rand_dm(rng, 4)
    base = DS.from_host(rho)
    batch = base.broadcast(3)
    gs = np.stack([rand_u(rng, 1) for _ in range(3)])
    batch.apply_gate_batched(gs, [0, 2, 3], [[1], [], [0, 1]])
    w = [0.2, 0.5, 0.3]
    exp = orc.ensemble(w, [orc.conjugate(orc.controlled_unitary(4, c, t, g_), rho) for g_, t, c in zip(gs, [0, 2, 3], [[1], [], [0, 1]])])
    assert close(np.asarray(batch.mix_branches(w)), exp)


def test_project_renorm(DS):
    rng = np.random.default_rng(10)
    psi = rand_ket(rng, 7)
    st = DS.from_host(psi)
    st.project_renorm([2, 5], 0b10)
    t = psi.reshape((2,) * 7).copy()
    mask = np.zeros((2,) * 7, dtype=bool)
    idx = [slice(None)] * 7
    idx[2], idx[5] = 1, 0
    mask[tuple(idx)] = True
    t[~mask] = 0
    t /= np.linalg.norm(t)
    assert close(np.asarray(st), t.reshape(-1))


def test_measure_matches_reference_cases(DS, golden):
    import qbot_b200
    from qbot_b200.host import hostmath as hm
    from qbot_b200.host.interp import Interpreter
    it = Interpreter(DS)
    bases = {'comp': hm.computation, 'hada': hm.hadamard, 'bell': hm.bell}
    for c in golden.cases('measure'):
        tg = c['targets_as_given']
        if tg is not None and c['targets_is_set']:
            tg = set(tg)
        r = it.state_ops['measure'](DS.from_host(golden.arr(c['rho'])), bases[c['basis']], tg, c['return_state'])
        assert np.allclose(r.probs, c['probs'], rtol=0, atol=2e-15), c
        assert list(r.basisSymbols) == c['symbols']
        assert close(np.asarray(r.unMeasuredDensity), golden.arr(c['unmeasured'])), c
        if c['return_state']:
            assert close(np.asarray(r.newState), golden.arr(c['new_state'])), c
    for c in golden.cases('replace'):
        out = it.state_ops['replace_arbitrary'](DS.from_host(golden.arr(c['rho'])), golden.arr(c['new']), c['targets'])
        assert close(np.asarray(out), golden.arr(c['out'])), c


@pytest.mark.parametrize('fusion', [False, True])
def test_rc_circuits_ket_and_dm(DS, golden, fusion):
    from qbot_b200.circuits import rc
    for key in [k for k in golden.rc.files if not k.endswith('_info')]:
        _, n, depth, seed = key.split('_')
        n, depth, seed = int(n), int(depth), int(seed)
        gates = rc(n, depth, seed)
        ket = DS.zero_state(n)
        ket.set_fusion(fusion)
        from qbot_b200 import DM
        dm = DS.zero_state(n, kind=DM)
        dm.set_fusion(fusion)
        for g in gates:
            ket.apply_gate(g.matrix(), g.target, g.controls)
            dm.apply_gate(g.matrix(), g.target, g.controls)
        assert close(np.asarray(dm), golden.rc[key], 1e-12), key          # reference rho
        assert close(np.asarray(ket.outer(True)), golden.rc[key], 1e-12), key   # rho_ref == psi psi^dagger
    # config-2 shape at parity size: rc(10, 200, 20) against the oracle ket path
    n = 10
    gates = rc(n, 200, 20)
    ket = DS.zero_state(n)
    ket.set_fusion(fusion)
    psi = np.zeros(1 << n, dtype=complex)
    psi[0] = 1
    for g in gates:
        ket.apply_gate(g.matrix(), g.target, g.controls)
        psi = orc.ket_apply(psi, n, g.target, g.matrix(), g.controls)
    assert close(np.asarray(ket), psi, 1e-12)


@pytest.mark.parametrize('fusion', [False, True])
def test_size_independent_properties_20q(DS, fusion):
    """Config 2 at full size (20 qubits, depth 200): properties that need no CPU reference --
    norm preservation and circuit followed by its inverse returning |0...0> exactly-ish."""
    from qbot_b200.circuits import rc
    n = 20
    gates = rc(n, 200, 20)
    st = DS.zero_state(n)
    st.set_fusion(fusion)
    for g in gates:
        st.apply_gate(g.matrix(), g.target, g.controls)
    assert abs(st.norm2()[0] - 1.0) < 1e-11
    p = st.probs([0, 5, 19])
    assert abs(p.sum() - 1.0) < 1e-11
    for g in reversed(gates):
        st.apply_gate(g.matrix().conj().T, g.target, g.controls)
    amp0 = st.download_range(0, 4)
    assert abs(amp0[0] - 1.0) < 1e-10 and np.max(np.abs(amp0[1:])) < 1e-10
    assert abs(st.probs([3])[0] - 1.0) < 1e-10
