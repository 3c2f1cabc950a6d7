"""CPU: host logic of the sharded ket (qubit map, exchange planning, gate localisation, probability
gather) on numpy shards -- P virtual ranks in threads, and world_size 2 over torch.distributed gloo.
The expected ket comes from the oracle's strided update on the full register."""
import os
import sys
import threading

import numpy as np
import pytest

from oracle import qbot_oracle as orc
from qbot_b200 import circuits
from qbot_b200.sharded import ShardedKet, QubitMap, make_lgate, select_pass
from np_shard import NumpyShard, VirtualComm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def extra_gates(n, rng):
    """2-qubit dense / diagonal / swap-like blocks so that multi-target localisation is covered."""
    out = []
    for _ in range(6):
        t = int(rng.integers(0, n - 1))
        kind = int(rng.integers(0, 3))
        if kind == 0:
            a = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
            m, _ = np.linalg.qr(a)
        elif kind == 1:
            m = np.diag(np.exp(1j * rng.uniform(0, 6.28, 4)))
        else:
            m = np.eye(4, dtype=complex)[[0, 2, 1, 3]]
        free = [q for q in range(n) if q not in (t, t + 1)]
        cs = [int(c) for c in rng.choice(free, size=int(rng.integers(0, 3)), replace=False)]
        out.append((m, t, cs))
    return out


def expected_ket(n, ops):
    psi = np.zeros(1 << n, dtype=complex)
    psi[0] = 1
    for m, t, cs in ops:
        psi = orc.ket_apply(psi, n, t, m, cs)
    return psi


def circuit_ops(n, depth, seed):
    ops = [(g.matrix(), g.target, list(g.controls)) for g in circuits.rc(n, depth, seed)]
    rng = np.random.default_rng(seed + 1)
    ex = extra_gates(n, rng)
    # interleave
    out = []
    step = max(1, len(ops) // (len(ex) + 1))
    for i, o in enumerate(ops):
        out.append(o)
        if i % step == step - 1 and ex:
            out.append(ex.pop())
    return out + ex


def run_virtual(n, world, ops, probe):
    shared = VirtualComm.Shared(world)
    results = [None] * world
    errors = []

    def work(rank):
        try:
            comm = VirtualComm(shared, rank)
            sk = ShardedKet(n, comm, shard_factory=NumpyShard)
            half = len(ops) // 2
            for m, t, cs in ops[:half]:
                sk.apply_gate(m, t, cs)
            sk.flush()
            for m, t, cs in ops[half:]:
                sk.apply_gate(m, t, cs)
            results[rank] = dict(ket=sk.gather(), probs=sk.probs(probe), norm=sk.norm2(),
                                 amps=sk.amplitudes([0, 5, (1 << n) - 1]), exchanges=sk.shard.exchanges,
                                 at=list(sk.map.at))
        except Exception as e:     # pragma: no cover
            errors.append(e)
            shared.barrier.abort()

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errors:
        raise errors[0]
    return results


@pytest.mark.parametrize('n,world', [(6, 2), (7, 4), (9, 8), (10, 4)])
def test_virtual_ranks_match_oracle(n, world):
    ops = circuit_ops(n, 6, 100 + n)
    want = expected_ket(n, ops)
    probe = [n - 1, 0, 2]
    res = run_virtual(n, world, ops, probe)
    for r in res:
        assert np.max(np.abs(r['ket'] - want)) < 1e-12
        assert np.max(np.abs(r['probs'] - orc.ket_probs(want, n, probe))) < 1e-12
        assert abs(r['norm'] - 1.0) < 1e-12
        assert np.max(np.abs(r['amps'] - want[[0, 5, (1 << n) - 1]])) < 1e-12
        assert r['at'] == res[0]['at']
    assert res[0]['exchanges'] >= 1          # the circuit does need global qubits


@pytest.mark.parametrize('n,world,split', [(15, 2, 2), (16, 4, 2), (16, 8, 1), (16, 4, 3)])
def test_pipelined_exchange_plan_matches_oracle(n, world, split):
    """Exchanges planned in 2^split pieces: parked bits on top of the shard (kept there from one exchange to the next
    when the gates around the exchange leave them alone, re-chosen otherwise), packed layout [parked][chunk][rest] --
    same ket as the oracle, and the pipelined path is actually taken."""
    ops = circuit_ops(n, 8, 300 + n)
    want = expected_ket(n, ops)
    shared = VirtualComm.Shared(world)
    out = [None] * world
    errors = []

    def work(rank):
        try:
            sk = ShardedKet(n, VirtualComm(shared, rank), shard_factory=NumpyShard, split=split)
            sk.min_first_phase = 6
            for rep in range(3):                 # later passes start from a permuted qubit map with parked bits on top
                for m, t, cs in ops:
                    sk.apply_gate(m, t, cs)
                sk.flush()
                if rep == 0:
                    first = sk.gather()
            out[rank] = dict(first=first, second=sk.gather(), split_exchanges=getattr(sk.shard, 'split_exchanges', 0),
                             send_side=getattr(sk.shard, 'send_side_exchanges', 0),
                             exchanges=sk.shard.exchanges, probs=sk.probs([0, n - 1]))
        except Exception as e:     # pragma: no cover
            errors.append(e)
            shared.barrier.abort()

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    if errors:
        raise errors[0]
    want2 = want
    for rep in range(2):
        for m, t, cs in ops:
            want2 = orc.ket_apply(want2, n, t, m, cs)
    for r in range(world):
        assert np.max(np.abs(out[r]['first'] - want)) < 1e-12
        assert np.max(np.abs(out[r]['second'] - want2)) < 1e-12
        assert np.max(np.abs(out[r]['probs'] - orc.ket_probs(want2, n, [0, n - 1]))) < 1e-12
        assert out[r]['split_exchanges'] >= 1, out[r]


def test_rank_scalar_under_controls_on_every_local_bit():
    """A diagonal gate whose target is a rank bit is a per-rank scalar on the amplitudes its controls select; with
    only two local bits and both of them controls there is no free local bit to carry diag(s, s) -- it becomes
    diag(1, s) on one of the controls."""
    n, world = 4, 4
    rz = circuits.z_rot(0.7)
    H = circuits.HADAMARD
    ops = [(H, q, []) for q in range(n)] + [(rz, 0, [2, 3]), (rz, 1, [3, 2]), (H, 3, []), (rz, 0, [3])]
    want = expected_ket(n, ops)
    for r in run_virtual(n, world, ops, [0, 3]):
        assert np.max(np.abs(r['ket'] - want)) < 1e-12
    mp = QubitMap(4, 2)
    m, tb, cm = mp.localise(make_lgate(rz, [3], [0, 1]), 2)      # logical bit 3 = rank bit 1, set on rank 2
    assert tb == [0] and cm == 0b10 and np.allclose(m, np.diag([1, rz[1, 1]]))


def test_reset_zero_gives_a_fresh_register():
    """reset_zero: |0...0> and the identity qubit map again (used by the multi-GPU e2e loop)"""
    n, world = 8, 4
    ops = circuit_ops(n, 5, 7)
    want = expected_ket(n, ops)
    shared = VirtualComm.Shared(world)
    out = [None] * world

    def work(rank):
        sk = ShardedKet(n, VirtualComm(shared, rank), shard_factory=NumpyShard)
        for rep in range(2):
            sk.reset_zero()
            assert sk.map.at == list(range(n))
            for m, t, cs in ops:
                sk.apply_gate(m, t, cs)
            out[rank] = sk.gather()

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for r in range(world):
        assert out[r] is not None and np.max(np.abs(out[r] - want)) < 1e-12


def test_exchange_count_is_small():
    # rc(12, 10): every layer writes ~2/3 of the qubits; farthest-next-use eviction keeps the
    # number of exchanges well below one per layer-and-rank-bit
    n, world = 12, 8
    ops = [(g.matrix(), g.target, list(g.controls)) for g in circuits.rc(n, 10, 34)]
    res = run_virtual(n, world, ops, [0])
    assert np.max(np.abs(res[0]['ket'] - expected_ket(n, ops))) < 1e-12
    assert res[0]['exchanges'] <= 12


def test_select_pass_and_localise_rules():
    H = circuits.HADAMARD
    X = circuits.PAULI_X
    rz = circuits.z_rot(0.3)
    g0 = make_lgate(H, [3])                 # writes 3
    g1 = make_lgate(rz, [3])                # reads 3
    g2 = make_lgate(X, [1], [3])            # writes 1, reads 3
    g3 = make_lgate(H, [0])
    assert g0.wmask == 8 and g1.wmask == 0 and g1.rmask == 8 and g2.wmask == 2 and g2.rmask == 8
    # bit 3 not writable: g0 stays, g1 and g2 touch bit 3 written by g0 -> blocked, g3 free
    assert select_pass([g0, g1, g2, g3], 0b0111) == [3]
    assert select_pass([g1, g2, g0, g3], 0b0111) == [0, 1, 3]
    mp = QubitMap(4, 1)                      # logical bit 3 selects the rank
    assert mp.localise(g2, 0) is None
    m, tb, cm = mp.localise(g2, 1)
    assert tb == [1] and cm == 0 and np.array_equal(m, X)
    m, tb, cm = mp.localise(g1, 1)           # per-rank scalar
    assert np.allclose(m, rz[1, 1] * np.eye(2))
    # a CZ written as a 4x4 matrix is block-diagonal in both targets
    cz = make_lgate(np.diag([1, 1, 1, -1]).astype(complex), [3, 2])
    assert cz.wmask == 0
    m, tb, cm = mp.localise(cz, 1)
    assert tb == [2] and np.allclose(m, np.diag([1, -1]))
    with pytest.raises(RuntimeError):
        mp.localise(g0, 0)


def _gloo_worker(rank, world, n, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch.distributed as dist
    from qbot_b200.sharded import ShardedKet, TorchComm
    from np_shard import NumpyShard
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        ops = circuit_ops(n, 5, 7)
        sk = ShardedKet(n, TorchComm(), shard_factory=NumpyShard)
        for m, t, cs in ops:
            sk.apply_gate(m, t, cs)
        ket = sk.gather()
        pr = sk.probs([0, n - 1])
        q.put((rank, ket, pr, sk.shard.exchanges))
    finally:
        dist.destroy_process_group()


def test_gloo_world_size_2():
    import torch.multiprocessing as mp
    n, world = 7, 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, n, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = expected_ket(n, circuit_ops(n, 5, 7))
    for rank, ket, pr, exch in got:
        assert np.max(np.abs(ket - want)) < 1e-12
        assert np.max(np.abs(pr - orc.ket_probs(want, n, [0, n - 1]))) < 1e-12
        assert exch >= 1


# ---------------------------------------------------------------------------------------------
# the sharded ket behind the DSL ops (qbot_b200/sharded_register.py): `qset` of a large product ket
# under one-process-per-rank gives a sharded register that gate / swap / peek drive
# ---------------------------------------------------------------------------------------------
SHARDED_PROGRAM_N = 16      # tensorExp(.., 14) is the smallest product the DSL keeps as a descriptor (hostmath.LAZY_MIN_QUBITS)


def sharded_program(n):
    gates = circuits.rc(n, 3, 5)
    lines = [f"qset tensorProd(hadamard.kets[0], tensorExp(comp.kets[0], {n - 2}), comp.kets[1])"]
    ops = []
    for i, g in enumerate(gates):
        lines.append(g.dsl())
        ops.append((g.matrix(), g.target, list(g.controls)))
        if i == len(gates) // 2:
            lines.append("swap 1 ; 9")
            ops.append(('swap', 1, 9))
    lines.append(f"peek r ; comp ; [0, 5, {n - 1}]")
    lines.append("peek b ; bell ; [3, 0]")
    lines.append(f"peek h ; hadamard ; [{n - 2}]")
    lines.append("cdef rho_a ; r.unMeasuredDensity")
    return "\n".join(lines) + "\n", ops


def expected_sharded_program(n, ops):
    plus = np.array([1, 1], dtype=complex) / np.sqrt(2)
    psi = np.array([1], dtype=complex)
    for f in [plus] + [np.array([1, 0], dtype=complex)] * (n - 2) + [np.array([0, 1], dtype=complex)]:
        psi = np.kron(psi, f)
    for op in ops:
        psi = orc.ket_swap(psi, n, op[1], op[2]) if isinstance(op[0], str) else orc.ket_apply(psi, n, op[1], op[0], op[2])
    return psi


DISC_DROP = [0, 2, 3, 5, 8, 9, 12, 15]


def _register_worker(rank, world, n, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch.distributed as dist
    import qbot_b200
    from qbot_b200 import sharded_register as sr
    from qbot_b200.sharded import TorchComm
    from np_shard import NumpyShard
    from fake_backend import FakeState
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        sr.enable(TorchComm(), shard_factory=NumpyShard, min_qubits=n)
        text, _ = sharded_program(n)
        ns = qbot_b200.executeTxt(text, state_cls=FakeState)
        reg = ns['state']
        out = dict(kind=type(reg).__name__, r=list(ns['r'].probs), b=list(ns['b'].probs), h=list(ns['h'].probs),
                   rho_a=np.asarray(ns['rho_a']), ket=np.asarray(reg), exchanges=reg.stats()['exchanges'])
        # a second program on a fresh interpreter reuses the shards of the dropped register
        del ns, reg
        ns2 = qbot_b200.executeTxt(f"qset tensorExp(comp.kets[0], {n})\ngate hadamardGate ; 0\ngate pauliXGate ; {n - 1} ; [0]\npeek p ; comp ; [0, {n - 1}]\n",
                                   state_cls=FakeState)
        out['p2'] = list(ns2['p'].probs)
        out['err'] = None
        try:
            qbot_b200.executeTxt(f"qset tensorExp(comp.kets[0], {n})\ngate hadamardGate ; 0 ; [] ; ProbVal([.5, .5], [True, False])\n", state_cls=FakeState)
        except SystemExit:
            out['err'] = 'exit'
        # `disc` on the sharded ket: Tr_rest psi psi^dagger of the kept qubits (made local, per-rank partial, one all-reduce)
        # becomes an ordinary density-matrix register; too many kept qubits are refused with the formatted error
        gl = "\n".join(g.dsl() for g in circuits.rc(n, 3, 5))
        ns3 = qbot_b200.executeTxt(f"qset tensorExp(comp.kets[0], {n})\n{gl}\ndisc {DISC_DROP}\npeek d ; comp ; [0, 3]\n", state_cls=FakeState)
        out['disc_rho'] = np.asarray(ns3['state'])
        out['disc_kind'] = type(ns3['state']).__name__
        out['disc_peek'] = list(ns3['d'].probs)
        out['disc_err'] = None
        try:
            import io
            from contextlib import redirect_stdout
            with redirect_stdout(io.StringIO()):
                qbot_b200.executeTxt(f"qset tensorExp(comp.kets[0], {n})\ndisc [0]\n", state_cls=FakeState)
        except SystemExit:
            out['disc_err'] = 'exit'
        # rho_A of a peek on a sharded ket is computed when it is first read: after an update of the register that read is
        # refused (formatted error) instead of handing out rho_A of another state; read at once, it is the peeked state's
        import io as _io
        from contextlib import redirect_stdout as _redirect
        head = f"qset tensorExp(hadamard.kets[0], {n})\npeek s ; comp ; [1, 2]\n"
        ns4 = qbot_b200.executeTxt(head + "cdef a ; np_array(s.unMeasuredDensity)\ngate pauliZGate ; 1\n", state_cls=FakeState)
        out['rho_at_once'] = np.asarray(ns4['a'])
        buf = _io.StringIO()
        out['stale'] = None
        try:
            with _redirect(buf):
                qbot_b200.executeTxt(head + "gate pauliZGate ; 1\ncdef a ; np_array(s.unMeasuredDensity)\n", state_cls=FakeState)
        except SystemExit:
            out['stale'] = buf.getvalue()
        q.put((rank, out))
    except BaseException as e:      # noqa: BLE001  (a formatted DSL error ends in sys.exit(): report it instead of hanging the parent)
        q.put((rank, dict(failed=f"{type(e).__name__}: {e}")))
        raise
    finally:
        sr.disable()
        dist.destroy_process_group()


def test_sharded_register_behind_the_dsl_ops_gloo():
    import torch.multiprocessing as mp
    n, world = SHARDED_PROGRAM_N, 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29900 + (os.getpid() % 90)
    procs = [ctx.Process(target=_register_worker, args=(r, world, n, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    _, ops = sharded_program(n)
    psi = expected_sharded_program(n, ops)
    from qbot_b200.host import hostmath as hm

    def weights(targets, basis):
        w = orc.basis_weights(psi, n, sorted(targets), basis.kets)
        w = w / w.sum()
        return [round(float(x / w.sum()), 15) for x in w]

    want_r = dict(probs=weights([0, 5, n - 1], hm.computation))
    want_b = dict(probs=weights([3, 0], hm.bell))
    want_h = dict(probs=weights([n - 2], hm.hadamard))
    t = psi.reshape([2] * n)
    keep = [0, 5, n - 1]
    mm = np.ascontiguousarray(t.transpose(keep + [a for a in range(n) if a not in keep])).reshape(8, -1)
    want_r['rho_a'] = mm @ mm.conj().T
    psi3 = np.zeros(1 << n, dtype=complex)
    psi3[0] = 1
    for g in circuits.rc(n, 3, 5):
        psi3 = orc.ket_apply(psi3, n, g.target, g.matrix(), g.controls)
    keep3 = [a for a in range(n) if a not in DISC_DROP]
    m3 = np.ascontiguousarray(psi3.reshape([2] * n).transpose(keep3 + DISC_DROP)).reshape(1 << len(keep3), -1)
    want_disc = m3 @ m3.conj().T
    d3 = np.real(np.diag(want_disc)).reshape([2] * len(keep3))
    want_disc_peek = d3.sum(axis=tuple(a for a in range(len(keep3)) if a not in (0, 3))).reshape(-1)
    for rank, out in got:
        assert 'failed' not in out, out
        assert out['kind'] == 'ShardedRegister'
        assert np.max(np.abs(out['ket'] - psi)) < 1e-12
        for key, want in (('r', want_r), ('b', want_b), ('h', want_h)):
            assert np.max(np.abs(np.array(out[key]) - np.array(want['probs']))) < 1e-12, key
        assert np.max(np.abs(out['rho_a'] - want_r['rho_a'])) < 1e-12
        assert out['p2'] == [0.5, 0.0, 0.0, 0.5]
        assert out['err'] == 'exit'          # a ProbVal condition leaves a mixed state: refused on a sharded ket, formatted error
        assert np.max(np.abs(out['rho_at_once'] - np.full((4, 4), 0.25))) < 1e-12
        assert out['stale'] is not None and 'has been updated since the peek' in out['stale']
        assert out['disc_kind'] != 'ShardedRegister' and out['disc_err'] == 'exit'
        assert np.max(np.abs(out['disc_rho'] - want_disc)) < 1e-12
        assert np.max(np.abs(np.array(out['disc_peek']) - want_disc_peek)) < 1e-12


# ---------------------------------------------------------------------------------------------
# a fresh product ket starts with the qubits of its choice on the rank bits (QubitMap.choose_initial)
# ---------------------------------------------------------------------------------------------
def _product_case(n, world, seed, lazy, two_programs=False):
    rng = np.random.default_rng(seed)
    factors = rng.normal(size=(n, 2)) + 1j * rng.normal(size=(n, 2))
    factors /= np.linalg.norm(factors, axis=1, keepdims=True)
    ops = circuit_ops(n, 5, seed)
    psi = np.array([1.0 + 0j])
    for q in range(n):
        psi = np.kron(psi, factors[q])
    want = psi
    for m, t, cs in ops:
        want = orc.ket_apply(want, n, t, m, cs)
    shared = VirtualComm.Shared(world)
    out, errors = [None] * world, []

    def work(rank):
        try:
            sk = ShardedKet(n, VirtualComm(shared, rank), shard_factory=NumpyShard)
            sk.lazy_map = lazy
            for rep in range(2 if two_programs else 1):
                sk.init_product(list(factors))
                for m, t, cs in ops:
                    sk.apply_gate(m, t, cs)
                ket = sk.gather()
            out[rank] = dict(ket=ket, at=list(sk.map.at), exchanges=sk.shard.exchanges, norm=sk.norm2())
        except Exception as e:     # pragma: no cover
            errors.append(e)
            shared.barrier.abort()

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errors:
        raise errors[0]
    return out, want


@pytest.mark.parametrize('n,world', [(6, 2), (8, 4), (9, 8)])
@pytest.mark.parametrize('lazy', [True, False])
def test_product_ket_with_chosen_initial_map_matches_oracle(n, world, lazy):
    out, want = _product_case(n, world, 500 + n, lazy)
    for r in out:
        assert np.max(np.abs(r['ket'] - want)) < 1e-12
        assert abs(r['norm'] - 1) < 1e-12
        assert r['at'] == out[0]['at']
    # a second program on the same (pooled) ket starts from a fresh map again
    out2, want2 = _product_case(n, world, 500 + n, lazy, two_programs=True)
    for r in out2:
        assert np.max(np.abs(r['ket'] - want2)) < 1e-12


def test_product_ket_initial_map_with_relabelled_qubits():
    """`swap` before the first flush only relabels the caller's qubits: the factors stay with the stored ket's qubits"""
    n, world = 8, 4
    # single-qubit circuits only, so that relabelled multi-qubit blocks stay contiguous in the check above
    rng = np.random.default_rng(9)
    factors = rng.normal(size=(n, 2)) + 1j * rng.normal(size=(n, 2))
    factors /= np.linalg.norm(factors, axis=1, keepdims=True)
    gates = [g for g in circuits.rc(n, 6, 77)]
    swaps = [(0, 5), (2, 7), (5, 1)]
    perm = list(range(n))
    for a, b in swaps:
        perm[a], perm[b] = perm[b], perm[a]
    psi = np.array([1.0 + 0j])
    for q in range(n):
        psi = np.kron(psi, factors[q])
    for g in gates:
        psi = orc.ket_apply(psi, n, perm[g.target], g.matrix(), [perm[c] for c in g.controls])
    want = np.ascontiguousarray(psi.reshape([2] * n).transpose(perm)).reshape(-1)
    shared = VirtualComm.Shared(world)
    out, errors = [None] * world, []

    def work(rank):
        try:
            sk = ShardedKet(n, VirtualComm(shared, rank), shard_factory=NumpyShard)
            sk.init_product(list(factors))
            for a, b in swaps:
                sk.swap_qubits(a, b)
            for g in gates:
                sk.apply_gate(g.matrix(), g.target, g.controls)
            out[rank] = sk.gather()
        except Exception as e:     # pragma: no cover
            errors.append(e)
            shared.barrier.abort()

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errors:
        raise errors[0]
    for k in out:
        assert np.max(np.abs(k - want)) < 1e-12


class _PlanOnlyShard:
    """records what a shard would be asked to do (no amplitudes): the exchange plan of a 34-qubit circuit"""
    supports_split = False

    def __init__(self, nl, comm):
        self.nl, self.exchanges, self.segments, self._n = nl, 0, [], 0

    def init_basis(self, has_one, local_index=0):
        pass

    def init_product(self, local_factors, coeff):
        assert len(local_factors) == self.nl

    def apply(self, m, tpos, cmask):
        assert all(0 <= p < self.nl for p in tpos) and cmask >> self.nl == 0
        self._n += 1

    def flush(self):
        if self._n:
            self.segments.append(self._n)
        self._n = 0

    def do_exchange(self, ex):
        self.exchanges += 1

    def sync(self):
        pass


class _OneRank:
    def __init__(self, rank, world):
        self.rank, self.world = rank, world

    def barrier(self):
        pass


def test_fresh_register_of_the_benchmark_needs_one_exchange():
    """BASELINE config 5 as a .qb program (what the multi-GPU e2e leg runs): rc(34, 10, 34) on a fresh product ket on
    8 ranks.  From the identity map the planner needs two exchanges (the second one for the last gate alone); with
    the initial map chosen from the queued gates it needs one."""
    n, world = 34, 8
    gates = circuits.rc(n, 10, n)
    zero = [np.array([1, 0], dtype=complex)] * n
    counts = {}
    for lazy in (False, True):
        sk = ShardedKet(n, _OneRank(3, world), shard_factory=_PlanOnlyShard)
        sk.lazy_map = lazy
        sk.init_product(zero)
        for g in gates:
            sk.apply_gate(g.matrix(), g.target, g.controls)
        sk.flush()
        counts[lazy] = (sk.shard.exchanges, list(sk.shard.segments))
    assert counts[False][0] == 2 and counts[True][0] == 1, counts
    assert sum(counts[True][1]) <= len(gates)


# ---------------------------------------------------------------------------------------------
# the whole host-visible chain on the CPU: product start with a chosen map -> local gate lists -> fusion planner ->
# the specialiser's generated kernel text (compiled with g++, tests/jit_emu.py) -> exchange -> ...
# ---------------------------------------------------------------------------------------------
class _GeneratedCodeShard(NumpyShard):
    """a numpy shard whose queued gates run as the fused sweeps the CUDA shard would launch: the same plan, the same
    generated source (CPU definitions of its macros), instead of one numpy update per gate"""

    def __init__(self, nl, comm):
        super().__init__(nl, comm)
        self.pending, self.sweeps = [], 0

    def apply(self, m, tpos, cmask):
        self.pending.append((np.array(m, dtype=complex), [int(p) for p in tpos], int(cmask)))
        self.applied += 1

    def flush(self):
        if self.pending:
            import jit_emu
            with _GeneratedCodeShard.lock:                       # one g++ at a time; the ranks are threads
                self.psi, st = jit_emu.run(self.nl, self.pending, self.psi)
            self.sweeps += st['sweeps']
            self.pending = []

    def do_exchange(self, ex):
        self.flush()
        super().do_exchange(ex)

    def download(self):
        self.flush()
        return super().download()

    def probs_local(self, positions):
        self.flush()
        return super().probs_local(positions)


_GeneratedCodeShard.lock = threading.Lock()


@pytest.mark.parametrize('n,world,depth,seed,lazy', [(16, 4, 8, 15, True), (15, 2, 12, 3, True), (16, 4, 8, 15, False)])
def test_fresh_product_register_through_planner_and_generated_kernels(n, world, depth, seed, lazy):
    # (14 / 13 local bits: free tile bits above the 12-bit tile)
    rng = np.random.default_rng(41)
    factors = rng.normal(size=(n, 2)) + 1j * rng.normal(size=(n, 2))
    factors /= np.linalg.norm(factors, axis=1, keepdims=True)
    gates = circuits.rc(n, depth, seed)
    want = np.array([1.0 + 0j])
    for q in range(n):
        want = np.kron(want, factors[q])
    for g in gates:
        want = orc.ket_apply(want, n, g.target, g.matrix(), g.controls)
    shared = VirtualComm.Shared(world)
    out, errors = [None] * world, []

    def work(rank):
        try:
            sk = ShardedKet(n, VirtualComm(shared, rank), shard_factory=_GeneratedCodeShard)
            sk.lazy_map = lazy
            sk.init_product(list(factors))
            for g in gates:
                sk.apply_gate(g.matrix(), g.target, g.controls)
            out[rank] = dict(ket=sk.gather(), probs=sk.probs([0, 7, n - 1]), sweeps=sk.shard.sweeps, exchanges=sk.shard.exchanges,
                             at=list(sk.map.at))
        except Exception as e:     # pragma: no cover
            errors.append(e)
            shared.barrier.abort()

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errors:
        raise errors[0]
    for r in out:
        assert np.max(np.abs(r['ket'] - want)) < 1e-12
        assert np.max(np.abs(r['probs'] - orc.ket_probs(want, n, [0, 7, n - 1]))) < 1e-12
        assert r['sweeps'] >= 1 and r['exchanges'] >= 1
    assert all(r['at'] == out[0]['at'] for r in out)


def test_requalified_register_reuses_the_shards():
    """a program that sets the (sharded) register again -- a loop that re-initialises it -- lets go of the old register before
    the new one is built, so the pooled shards carry it: one ShardedKet for the whole program, not two"""
    import qbot_b200
    from qbot_b200 import sharded_register as sr
    from fake_backend import FakeState
    made = []

    class CountingShard(NumpyShard):
        def __init__(self, nl, comm):
            super().__init__(nl, comm)
            made.append(self)

    n = 14
    sr.enable(VirtualComm(VirtualComm.Shared(1), 0), shard_factory=CountingShard, min_qubits=n)
    try:
        prog = "\n".join(["cdef i ; 0", "mark loop", f"qset tensorExp(hadamard.kets[0], {n})", "gate pauliZGate ; i",
                          "peek p ; hadamard ; [0, 1, 2]", "cdef i ; i + 1", "cjmp loop ; i < 3"]) + "\n"
        ns = qbot_b200.executeTxt(prog, state_cls=FakeState)
        assert type(ns['state']).__name__ == 'ShardedRegister'
        # the last iteration flipped qubit 2 from |+> to |->: outcome (0, 0, 1) in the hadamard basis
        assert np.allclose(ns['p'].probs, [0, 1, 0, 0, 0, 0, 0, 0], atol=1e-12)
        assert len(made) == 1, len(made)
        del ns
    finally:
        sr.disable()
