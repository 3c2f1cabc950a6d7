"""Host logic (interpreter mirror, ops, ProbVal) against programs run through the REAL
reference (tests/golden/scripts.json), with the numpy test double standing in for the device.
CPU only; the same programs run against the CUDA backend in test_gpu_scripts.py."""
import io
import json
from contextlib import redirect_stdout

import numpy as np
import pytest

import qbot_b200
from qbot_b200.host.probval import ProbVal, funcWrapper
from qbot_b200.host import hostmath as hm
from qbot_b200.host.ops import GateDesc, SwapDesc
from fake_backend import FakeState
from conftest import close


def run_script(text, state_cls):
    buf = io.StringIO()
    exited = False
    ns = {}
    try:
        with redirect_stdout(buf):
            ns = qbot_b200.executeTxt(text, state_cls=state_cls)
    except SystemExit:
        exited = True
    return ns, buf.getvalue(), exited


def check_script(rec, arrays, state_cls, rtol=1e-12, prob_tol=0.0):
    ns, out, exited = run_script(rec['text'], state_cls)
    assert exited == rec['exited'], rec['name']
    assert out == rec['stdout'], rec['name']
    if 'state' in rec:
        assert close(np.asarray(ns['state']), arrays[rec['state']], rtol), rec['name']
    for name, exp in rec['vars'].items():
        got = ns[name]
        if exp['type'] == 'meas':
            if prob_tol:        # device arithmetic: same outcomes in the same order, weights to the fp64 tolerance
                assert len(got.probs) == len(exp['probs']) and np.allclose(got.probs, exp['probs'], rtol=0, atol=prob_tol), (rec['name'], got.probs)
            else:
                assert got.probs == exp['probs'], (rec['name'], got.probs)
            assert list(got.basisSymbols) == exp['symbols']
            assert close(np.asarray(got.unMeasuredDensity), arrays[f"{rec['state']}_{name}_un"], rtol)
        elif exp['type'] == 'array':
            assert close(np.asarray(got), arrays[exp['key']], rtol)
        elif exp['type'] == 'py':
            if isinstance(exp['value'], float):          # a number read out of the register by a user expression
                assert abs(float(got) - exp['value']) <= max(prob_tol, 1e-15), (rec['name'], name, got)
            else:
                assert json.loads(json.dumps(got, default=str)) == exp['value'], rec['name']


def test_all_golden_scripts(golden):
    assert len(golden.scripts) >= 60
    for rec in golden.scripts:
        check_script(rec, golden.scripts_arr, FakeState)


def test_fuzzed_scripts(golden):
    """random programs (scripts/fuzz_dsl.py: every state op with plain and ProbVal arguments, conditions,
    sub-register qset, disc, meas / peek in three bases, injected errors) recorded from the real reference"""
    assert len(golden.scripts_fuzz) >= 150
    assert sum(r['exited'] for r in golden.scripts_fuzz) >= 20
    for rec in golden.scripts_fuzz:
        check_script(rec, golden.scripts_fuzz_arr, FakeState)


def test_fuzzed_scripts_on_5_and_6_qubits(golden):
    """the same generator on 5-6 qubit registers, inside what the reference still gets right at that size (no swap,
    slot-aligned controls only: SURVEY.md F5 / F6), recorded from the real reference"""
    assert len(golden.scripts_fuzz_big) >= 60
    for rec in golden.scripts_fuzz_big:
        check_script(rec, golden.scripts_fuzz_big_arr, FakeState)


def test_probval_rules(golden):
    for c in golden.probval:
        if c['kind'] == 'normalize':
            vals = [tuple(v) for v in c['values']] if c['tuple_values'] else c['values']
            pv = ProbVal(c['probs'], vals)
            assert pv.probs == c['out_probs']
            assert [list(v) if isinstance(v, tuple) else v for v in pv.values] == c['out_values']
        elif c['kind'] == 'nested':
            pv = ProbVal([.25, .75], [ProbVal([.5, .5], [10, 20]), 30])
            assert pv.probs == c['out_probs'] and pv.values == c['out_values']
        elif c['kind'] == 'unwrap':
            assert ProbVal.fromUnzipped([.5, .5], [7, 7]) == c['result']
    a = ProbVal([.5, .5], [0, 1])
    b = ProbVal([.2, .3, .5], [10, 20, 30])
    by_kind = {(c['kind'], c.get('expr')): c for c in golden.probval}
    r = funcWrapper(lambda x, c, y: (x, c, y), a, 'k', b)
    assert [list(v) for v in r.values] == by_kind[('fanout', None)]['out_values'] and r.probs == by_kind[('fanout', None)]['out_probs']
    r2 = funcWrapper(lambda x, y: x + y, a, b)
    assert r2.values == by_kind[('fanout_sum', None)]['out_values'] and r2.probs == by_kind[('fanout_sum', None)]['out_probs']
    for expr, val in (('a*2+1', a * 2 + 1), ('a+b', a + b), ('a==0', a == 0), ('-b', -b), ('a-b', a - b), ('3-a', 3 - a)):
        exp = by_kind.get(('arith', expr))
        if exp is not None:
            assert val.probs == exp['out_probs'] and val.values == exp['out_values'], expr


def test_gate_descriptor_equality_matches_full_unitaries():
    """GateDesc.__eq__ must agree with elementwise equality of the unitaries the reference
    would have built (it is what ProbVal.normalize de-duplicates on)."""
    from oracle import qbot_oracle as orc
    H = hm.tensor_prod(np.eye(2), np.array([[1, 1], [1, -1]]) / np.sqrt(2))
    X = np.array([[0, 1], [1, 0]], dtype=complex)
    Z = np.diag([1, -1]).astype(complex)
    I = np.eye(2, dtype=complex)
    n = 4
    descs = [GateDesc(X, 1, []), GateDesc(np.kron(I, X), 0, []), GateDesc(np.kron(X, I), 1, []), GateDesc(X, 1, [0]),
             GateDesc(X, 1, [0, 3]), GateDesc(X, 1, [3, 0]), GateDesc(Z, 1, [2]), GateDesc(Z, 2, [1]), GateDesc(I, 0, [1]),
             GateDesc(np.eye(4), 2, []), GateDesc(X, 2, []), GateDesc(np.kron(I, np.kron(X, I)), 0, [])]
    full = [orc.controlled_unitary(n, d.controls, d.target, d.matrix) for d in descs]
    for i, a in enumerate(descs):
        for j, b in enumerate(descs):
            assert (a == b) == bool((full[i] == full[j]).all()), (i, j)
    assert SwapDesc(1, 2) == SwapDesc(2, 1) and SwapDesc(0, 0) == SwapDesc(3, 3) and not (SwapDesc(0, 1) == SwapDesc(0, 2))


def test_inplace_only_when_unaliased():
    # `cdef old ; state` aliases the register: the next gate must not change `old`
    ns, _, _ = run_script("qset comp[0]\ncdef old ; state\ngate pauliXGate ; 0\n", FakeState)
    assert np.asarray(ns['old'])[0, 0] == 1 and np.asarray(ns['state'])[1, 1] == 1
    ns, _, _ = run_script("qset tensorExp(comp[0], 2)\nmeas x ; comp ; 0\ngate pauliXGate ; 1\n", FakeState)
    assert np.asarray(ns['x'].newState)[0, 0] == 1 and np.asarray(ns['state'])[1, 1] == 1


def test_ownership_is_explicit_not_refcounted():
    """in place only when the ops made the register themselves AND no expression names `state`"""
    from qbot_b200.host import ops as ops_mod
    lines_plain = ["qset comp[0]", "gate pauliXGate ; 0"]
    lines_named = ["qset comp[0]", "note state is only mentioned in a comment", "pydo l.append(state)"]
    assert not ops_mod._program_names_state({}, lines_plain)
    assert not ops_mod._program_names_state({}, lines_plain + ["note a comment about the state"])
    assert ops_mod._program_names_state({}, lines_named)
    # a register that came out of a user expression (here: the measurement result's newState) is shared
    ns, _, _ = run_script("qset tensorExp(comp[0], 2)\npeek x ; comp ; 0\nmeas y ; comp ; 0\nqset y.newState\ngate pauliXGate ; 1\n", FakeState)
    assert np.asarray(ns['y'].newState)[0, 0] == 1 and np.asarray(ns['state'])[1, 1] == 1
    # extra Python references held by somebody else (what broke the getrefcount rule) change nothing
    ns, _, _ = run_script("qset comp[0]\ngate pauliXGate ; 0\ngate hadamardGate ; 0\n", FakeState)
    assert not ns['state']._shared


def test_probval_gate_on_a_ket_register_gives_the_mixed_state():
    """a ProbVal-valued gate / condition / target on a ket-mode register must give
    sum_i p_i U_i rho U_i^dagger, i.e. the same as on the density register -- never sum_i p_i U_i psi"""
    ket = "qset np_array([1, 0, 0, 0]) * (1+0j)\n"
    dm = "qset tensorExp(comp[0], 2)\n"
    for body in ("gate pauliXGate ; 0 ; [] ; ProbVal([.5, .5], [True, False])\n",
                 "gate hadamardGate ; ProbVal([.25, .75], [0, 1])\n",
                 "gate hadamardGate ; 0\ngate pauliXGate ; 1 ; ProbVal([.5, .5], [[0], []])\n",
                 "gate ProbVal([.3, .7], [pauliXGate, hadamardGate]) ; 1\n",
                 "gate hadamardGate ; 0\nswap ProbVal([.5, .5], [0, 1]) ; 1\n"):
        a, _, _ = run_script(ket + body, FakeState)
        b, _, _ = run_script(dm + body, FakeState)
        ra, rb = np.asarray(a['state']), np.asarray(b['state'])
        assert ra.shape == (4, 4) and close(ra, rb), body
        assert abs(np.trace(ra) - 1) < 1e-14


def test_qset_with_repeated_targets_fails_like_the_reference():
    """`qset rho2 ; [1, 1]` passes the count check, then the reference dies in its
    permutation matmul with a ValueError (density.py:203-225)"""
    ns, out, exited = run_script("qset tensorExp(comp[0], 3)\nqset tensorExp(comp[1], 2) ; [1, 1]\n", FakeState)
    assert exited and "matmul: Input operand 1 has a mismatch in its core dimension 0" in out and "(size 16 is different from 8)" in out


def test_ket_register_new_representation():
    # a 1-D qset starts a ket-mode register (the reference has no working ket path)
    ns, _, _ = run_script("qset np_array([1, 0, 0, 0]) * (1+0j)\ngate hadamardGate ; 0\ngate pauliXGate ; 1 ; [0]\n", FakeState)
    psi = np.asarray(ns['state'])
    assert psi.shape == (4,) and close(psi, np.array([1, 0, 0, 1]) / np.sqrt(2))
    ns, _, _ = run_script("qset np_array([1, 0, 0, 0]) * (1+0j)\ngate hadamardGate ; 0\ngate pauliXGate ; 1 ; [0]\nmeas x ; bell\n", FakeState)
    assert ns['x'].probs == [1.0, 0.0, 0.0, 0.0]


def test_lazy_product_states_and_ket_peek():
    """tensorProd / tensorExp results of >= 14 qubits stay descriptors and are
    built by the device-side constructor; a large ket-mode register can be `peek`ed (outcome
    weights from the amplitudes, rho_A on demand) while `meas` -- whose reference collapse is a
    mixed product state -- is refused with the reference's error formatting."""
    import qbot_b200
    from qbot_b200.host import hostmath as hm
    from fake_backend import FakeState
    lp = hm.tensor_exp(np.array([1, 0], dtype=complex), 16)
    assert hm.is_lazy(lp) and lp.shape == (1 << 16,) and lp.ndim == 1 and lp.size == 1 << 16
    assert not hm.is_lazy(hm.tensor_exp(np.array([1, 0], dtype=complex), 13))
    mixed = hm.tensor_prod(hm.tensor_exp(np.array([1, 0], dtype=complex), 14), np.array([0, 1], dtype=complex))
    assert hm.is_lazy(mixed) and len(mixed.factors) == 15 and np.asarray(mixed)[1] == 1
    n = 14
    prog = "\n".join([f"qset tensorExp(comp.kets[0], {n})", "gate hadamardGate ; 3", "gate pauliXGate ; 9 ; [3]",
                      "gate zRotGate(0.7) ; 9", "peek r ; comp ; [9, 3]", "cdef rhoA ; r.unMeasuredDensity"])
    ns = qbot_b200.executeTxt(prog, state_cls=FakeState)
    assert ns['state'].kind == 0 and ns['state'].shape == (1 << n,)
    assert ns['r'].probs == [0.5, 0.0, 0.0, 0.5] and ns['r'].newState is None
    rho = np.asarray(ns['rhoA'])
    assert rho.shape == (4, 4) and abs(np.trace(rho) - 1) < 1e-14 and abs(abs(rho[0, 3]) - 0.5) < 1e-14
    with pytest.raises(SystemExit):
        qbot_b200.executeTxt(prog + "\nmeas m ; comp ; [0]", state_cls=FakeState)


def test_probval_normalize_fast_path_equals_pairwise_loop():
    """row f4: for exact-equality value kinds normalize uses hashing (O(B)); it must keep exactly
    what the reference's pairwise loop keeps -- first occurrence wins, later duplicates are dropped
    (their probability is NOT added), entries below 1e-5 are dropped when reached"""
    from qbot_b200.host import probval as pv
    from qbot_b200.host.ops import GateDesc
    rng = np.random.default_rng(3)

    def slow(probs, values):
        p, v = list(probs), list(values)
        i = 0
        while i < len(p):
            if p[i] < pv.smallVal:
                del p[i], v[i]
                continue
            j = i + 1
            while j < len(p):
                if pv.valsClose(v[i], v[j]):
                    del p[j], v[j]
                else:
                    j += 1
            i += 1
        t = sum(p)
        return [round(x / t, pv.probRounding) for x in p], v

    H = np.array([[1, 1], [1, -1]], dtype=complex) * 2 ** -0.5
    X = np.array([[0, 1], [1, 0]], dtype=complex)
    pools = {
        'ints': [int(x) for x in rng.integers(0, 12, 200)],
        'arrays': [np.array([[1, 0], [0, np.exp(1j * float(k))]]) for k in rng.integers(0, 9, 150)] + [np.array([[0.0, -0.0], [0, 1]]), np.array([[-0.0, 0.0], [0, 1]])] * 20,
        'gates': [GateDesc(H if k % 2 else X, int(k) % 5, [7] if k % 3 == 0 else []) for k in rng.integers(0, 30, 120)],
    }
    for name, vals in pools.items():
        probs = list(rng.uniform(0, 1, len(vals)))
        for k in rng.integers(0, len(vals), 15):
            probs[int(k)] = 1e-7                                     # entries below smallVal
        assert pv.ProbVal._exact_keys(vals) is not None, name
        got = pv.ProbVal(list(probs), list(vals))
        wp, wv = slow(probs, vals)
        assert got.probs == wp, name
        assert len(got.values) == len(wv) and all(pv.valsClose(a, b) for a, b in zip(got.values, wv)), name
    # floats compare with a tolerance: never the fast path
    assert pv.ProbVal._exact_keys([0.1 * k for k in range(40)]) is None
    # the 4096-branch case of config 4: descriptors with many duplicates, well under a second
    import time
    vals = [GateDesc(np.diag([1, np.exp(1j * (k % 64))]), k % 16, []) for k in range(4096)]
    t0 = time.perf_counter()
    r = pv.ProbVal([1.0 / 4096] * 4096, vals)
    assert len(r.values) == 64 * 16 // np.gcd(64, 16) or len(r.values) <= 1024
    assert time.perf_counter() - t0 < 5.0


def test_probval_measurement_targets_behind_the_flag(monkeypatch):
    """row f4 / F8: off -> the reference's crash text; on -> the mixture over the target sets, checked
    against the oracle's measurement on each branch"""
    from oracle import qbot_oracle as orc
    prog = ("qset tensorProd(hada[0], comp[1], comp[0])\ngate pauliXGate ; 2 ; [0]\n"
            "meas m ; comp ; ProbVal([.25, .75], [[0], [1]])\n")
    monkeypatch.delenv('QBOT_B200_PROBVAL_MEAS', raising=False)
    ns, out, exited = run_script(prog, FakeState)
    assert exited and "has no attribute 'probs'" in out
    monkeypatch.setenv('QBOT_B200_PROBVAL_MEAS', '1')
    ns, out, exited = run_script(prog, FakeState)
    assert not exited
    plus = np.array([1, 1]) / np.sqrt(2)
    psi = np.kron(np.kron(plus, [0, 1]), [1, 0]).astype(complex)
    rho = orc.dm_apply(np.outer(psi, psi.conj()), 3, 2, np.array([[0, 1], [1, 0]], dtype=complex), [0])
    comp = [np.diag([1, 0]).astype(complex), np.diag([0, 1]).astype(complex)]
    r0, r1 = orc.measure(rho, comp, [0], True), orc.measure(rho, comp, [1], True)
    want_p = [.25 * a + .75 * b for a, b in zip(r0['probs'], r1['probs'])]
    assert np.allclose(ns['m'].probs, want_p, atol=1e-15)
    assert close(np.asarray(ns['state']), .25 * r0['newState'] + .75 * r1['newState'])
    assert close(np.asarray(ns['m'].unMeasuredDensity), .25 * r0['unMeasuredDensity'] + .75 * r1['unMeasuredDensity'])


def _disc_ket_case(n, drop):
    from qbot_b200.circuits import rc
    from oracle import qbot_oracle as orc
    gates = rc(n, 3, 5)
    prog = "\n".join([f"qset tensorExp(comp.kets[0], {n})"] + [g.dsl() for g in gates] + [f"disc {drop}", "peek p ; comp ; [0, 3]"])
    psi = np.zeros(1 << n, dtype=complex)
    psi[0] = 1
    for g in gates:
        psi = orc.ket_apply(psi, n, g.target, g.matrix(), g.controls)
    keep = [q for q in range(n) if q not in drop]
    m = np.ascontiguousarray(psi.reshape([2] * n).transpose(keep + list(drop))).reshape(1 << len(keep), -1)
    rho = m @ m.conj().T
    d = np.real(np.diag(rho)).reshape([2] * len(keep))
    peek = d.sum(axis=tuple(a for a in range(len(keep)) if a not in (0, 3))).reshape(-1)
    return prog, rho, peek


def test_disc_on_a_large_ket_register():
    """`disc` on a ket-mode register of more than 13 qubits (the new representation): what is left is
    Tr_rest psi psi^dagger of the kept qubits, computed from the amplitudes -- an ordinary density-matrix register (the
    reference's disc, operators.py:169-188 -> density.partialTraceArbitrary, on a state it could never hold); a discard
    that would leave more than 13 qubits is refused with the formatted error."""
    n, drop = 16, [0, 2, 3, 5, 8, 9, 12, 15]
    prog, rho, peek = _disc_ket_case(n, drop)
    ns = qbot_b200.executeTxt(prog, state_cls=FakeState)
    st = ns['state']
    assert st.kind == 1 and st.nq == n - len(drop)
    assert np.max(np.abs(np.asarray(st) - rho)) < 1e-12
    assert np.max(np.abs(np.array(ns['p'].probs) - peek)) < 1e-12
    buf = io.StringIO()
    with pytest.raises(SystemExit), redirect_stdout(buf):
        qbot_b200.executeTxt(f"qset tensorExp(comp.kets[0], {n})\ndisc [0, 1]\n", state_cls=FakeState)
    assert "ket-mode register can be cut down to at most 13 qubits" in buf.getvalue()
    # small ket-mode registers keep going through psi psi^dagger (unchanged path)
    ns = qbot_b200.executeTxt("qset tensorExp(comp.kets[0], 14)\ngate hadamardGate ; 0\ngate pauliXGate ; 13 ; [0]\ndisc [13]\npeek p ; comp ; [0]\n",
                              state_cls=FakeState)
    assert ns['state'].nq == 13 and list(ns['p'].probs) == [0.5, 0.5]


def test_every_op_on_a_large_ket_register_answers_or_refuses_with_a_formatted_error():
    """A ket-mode register above 13 qubits has no 4^n form: every op either works on the amplitudes or ends like any
    other DSL error (formatted window + sys.exit, qbot/errors.py) -- never with a raw exception or a 64 GiB allocation."""
    n = 16
    head = f"qset tensorExp(comp.kets[0], {n})\ngate hadamardGate ; 0\n"
    works = {
        'peek_all': ("peek p ; comp\n", None),
        'peek_hada': ("peek p ; hadamard ; [0, 1]\n", [0.5, 0.5, 0.0, 0.0]),
        'peek_bell': ("peek p ; bell ; [0, 1]\n", [0.25, 0.25, 0.25, 0.25]),
        'swap': ("swap 0 ; 3\npeek p ; comp ; [3]\n", [0.5, 0.5]),
        'qset_again': (f"qset tensorExp(hadamard.kets[0], {n})\npeek p ; comp ; [0]\n", [0.5, 0.5]),
        'names_state': ("cdef x ; state\ngate hadamardGate ; 1\npeek p ; comp ; [1]\n", [0.5, 0.5]),
        'dense_block': ("gate qftGate(3) ; 2\npeek p ; comp ; [2]\n", [0.5, 0.5]),
        'disc_to_13': ("disc [0, 1, 2]\npeek p ; comp ; [0]\n", [1.0, 0.0]),
        'disc_all': (f"disc {list(range(n))}\n", None),
        'disc_probval': ("disc ProbVal([.5,.5],[[0,1,2],[3,4,5]])\n", None),
    }
    for name, (tail, probs) in works.items():
        ns = qbot_b200.executeTxt(head + tail, state_cls=FakeState)
        if probs is not None:
            assert np.allclose(ns['p'].probs, probs, atol=1e-12), name
    refused = {
        'meas': ("meas m ; comp ; [0]\n", "meas on a 16-qubit ket-mode register"),
        'disc_too_many_left': ("disc [0]\n", "at most 13 qubits"),
        'qset_sub_register': ("qset comp.kets[1] ; [0]\n", "cannot become a density matrix"),
        'probval_gate': ("gate ProbVal([.5,.5],[hadamardGate, pauliXGate]) ; 1\n", "leaves a mixed state"),
        'probval_target': ("gate hadamardGate ; ProbVal([.5,.5],[1,2])\n", "leaves a mixed state"),
        'probval_condition': ("gate hadamardGate ; 1 ; [] ; ProbVal([.5,.5],[True, False])\n", "leaves a mixed state"),
        'probval_swap': ("swap ProbVal([.5,.5],[0,1]) ; 3\n", "leaves a mixed state"),
        'range': (f"peek p ; comp ; [{n}]\n", "outside of valid range"),
        'bell_odd': ("peek p ; bell ; [0]\n", "must be divisable"),
    }
    for name, (tail, text) in refused.items():
        buf = io.StringIO()
        with pytest.raises(SystemExit), redirect_stdout(buf):
            qbot_b200.executeTxt(head + tail, state_cls=FakeState)
        assert text in buf.getvalue(), (name, buf.getvalue()[:300])


def test_single_process_programs_do_not_import_torch():
    """the sharded-register hook is consulted by every `qset` of a large product ket; without a launcher (WORLD_SIZE < 2,
    torch.distributed not loaded) it must answer without importing torch -- seconds on a cold start, and the single-GPU
    path needs nothing from it"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path[:0] = [%r, %r]\n"
            "import qbot_b200\nfrom fake_backend import FakeState\n"
            "ns = qbot_b200.executeTxt('qset tensorExp(comp.kets[0], 16)\\ngate hadamardGate ; 0\\npeek p ; comp ; [0]\\n', state_cls=FakeState)\n"
            "assert list(ns['p'].probs) == [0.5, 0.5]\n"
            "assert 'torch' not in sys.modules, 'torch was imported'\n") % (root, os.path.join(root, 'tests'))
    env = {k: v for k, v in os.environ.items() if k not in ('WORLD_SIZE', 'RANK', 'LOCAL_RANK', 'QBOT_B200_SHARD')}
    p = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300, env=env)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
