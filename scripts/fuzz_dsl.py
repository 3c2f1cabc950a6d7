"""Differential fuzzer of the host side (build container only, no GPU): random `.qb` programs
run through the REAL reference (`/root/reference`, executeTxt) and through this repo's interpreter
mirror + ops (`qbot_b200.executeTxt`) over the numpy test double; stdout, exit behaviour, the final
register and every named result must agree (state 1e-12, ProbVal ordering exact, outcome weights exact up to the 15th decimal place --
see probs_agree).

    PYTHONDONTWRITEBYTECODE=1 python scripts/fuzz_dsl.py --seeds 0:500 [--emit tests/golden/scripts_fuzz_src.json]

Programs stay inside the reference's validity domain: n <= 4 qubits, at most one
control unless the controls are slot-aligned, no ProbVal measurement targets.  `--emit` writes the
programs of the run in the `scripts_src.json` format so that `tests/golden/make_golden.py`-style
fixtures can be recorded from them.  TEST INFRASTRUCTURE."""
import argparse
import io
import json
import os
import re
import sys
from contextlib import redirect_stdout

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.dont_write_bytecode = True
sys.path.insert(0, '/root/reference')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

ONE_Q = ['comp[0]', 'comp[1]', 'hada[0]', 'hada[1]']
GATES1 = ['hadamardGate', 'pauliXGate', 'pauliYGate', 'pauliZGate', 'xRotGate(%.3f)', 'yRotGate(%.3f)', 'zRotGate(%.3f)']


def pv(rng, vals):
    p = rng.uniform(0.1, 1.0, len(vals))
    p = np.round(p / p.sum(), 3)
    p[-1] = round(1.0 - p[:-1].sum(), 3)
    return 'ProbVal([%s], [%s])' % (', '.join('%.3f' % x for x in p), ', '.join(vals))


def gate1(rng):
    g = GATES1[int(rng.integers(len(GATES1)))]
    return g % rng.uniform(0, 6.28) if '%' in g else g


def initial(rng, n):
    parts = []
    left = n
    while left:
        if left >= 2 and rng.random() < 0.3:
            parts.append('bell[%d]' % rng.integers(4))
            left -= 2
        else:
            parts.append(ONE_Q[int(rng.integers(4))])
            left -= 1
    order = rng.permutation(len(parts))
    parts = [parts[i] for i in order]
    return 'qset ' + (parts[0] if len(parts) == 1 else 'tensorProd(%s)' % ', '.join(parts))


def slot_aligned(n, t, cs):
    c = len(cs)
    return cs == (list(range(t - c, t)) if t > 0 else list(range(n - c, n)))


def line(rng, n, names, extra=False):
    r = rng.random()
    if extra and rng.random() < 0.2:               # forms kept out of the recorded fixtures (they print floats / reuse results)
        ms = [v for v in names if v.startswith('m')]
        q = rng.random()
        if ms and q < 0.35:
            return 'cout %s' % ms[int(rng.integers(len(ms)))]
        if ms and q < 0.6:
            m = ms[int(rng.integers(len(ms)))]
            return 'gate %s ; %d ; [] ; ProbVal(%s.probs, [k %% 2 == 0 for k in range(len(%s.probs))])' % (gate1(rng), rng.integers(n), m, m)
        if q < 0.75:
            return 'qset %s ; %d' % (pv(rng, ['comp[1]', 'hada[1]', 'comp[1]']), rng.integers(n))
        if q < 0.9:
            name = 'c%d' % len(names)
            names.append(name)
            return 'cdef %s ; %s * 2 + 1' % (name, pv(rng, ['1', '2', '1']))
        return 'gate %s ; %s ; [] ; %s' % (pv(rng, [gate1(rng), gate1(rng)]), pv(rng, [str(int(rng.integers(n))), str(int(rng.integers(n)))]), pv(rng, ['True', 'False']))
    if r < 0.40:                                   # gate, 0-2 controls, optional ProbVal pieces / condition
        k = 2 if (n >= 2 and rng.random() < 0.15) else 1
        t = int(rng.integers(0, n - k + 1))
        if k == 2:
            g = rng.choice(['qftGate(2)', 'tensorProd(hadamardGate, pauliXGate)', 'simonsGate(2, lambda x: x)'])
        else:
            g = gate1(rng)
        free = [q for q in range(n) if not t <= q < t + k]
        cs = []
        if free and rng.random() < 0.5:
            if len(free) >= 2 and rng.random() < 0.3:
                c = 2
                cand = list(range(t - c, t)) if t - c >= 0 else (list(range(n - c, n)) if t == 0 and k + c <= n else [])
                cs = cand if cand and slot_aligned(n, t, cand) and all(q in free for q in cand) else [int(rng.choice(free))]
            else:
                cs = [int(rng.choice(free))]
        gs, ts, css = g, str(t), str(cs)
        q = rng.random()
        if q < 0.12 and k == 1:
            gs = pv(rng, [gate1(rng), gate1(rng)])
        elif q < 0.22 and k == 1 and not cs and n >= 2:
            a, b = rng.choice(n, 2, replace=False)
            ts = pv(rng, [str(int(a)), str(int(b))])
        elif q < 0.30 and len(cs) == 1 and len(free) >= 2:
            a, b = rng.choice(free, 2, replace=False)
            css = pv(rng, [str(int(a)), str(int(b))])
        out = 'gate %s ; %s ; %s' % (gs, ts, css)
        q = rng.random()
        if q < 0.08:
            out += ' ; ' + pv(rng, ['True', 'False'])
        elif q < 0.14:
            out += ' ; %d > 1' % rng.integers(0, 4)
        return out
    if r < 0.50 and n >= 2:                        # swap
        a, b = rng.choice(n, 2, replace=False)
        if n >= 3 and rng.random() < 0.2:
            others = [q for q in range(n) if q != a]
            x, y = rng.choice(others, 2, replace=False)
            return 'swap %d ; %s' % (a, pv(rng, [str(int(x)), str(int(y))]))
        return 'swap %d ; %d' % (a, b)
    if r < 0.68:                                   # meas / peek
        basis = 'comp' if rng.random() < 0.6 else 'hada'
        m = int(rng.integers(1, n + 1))
        tg = sorted(int(x) for x in rng.choice(n, m, replace=False))
        if rng.random() < 0.15 and n >= 2:
            basis, m = 'bell', 2
            tg = sorted(int(x) for x in rng.choice(n, 2, replace=False))
        name = 'm%d' % len(names)
        names.append(name)
        form = rng.choice(['%s', '(%s,)', '{%s}']) if m == 1 else rng.choice(['[%s]', '(%s)', '{%s}'])
        tgs = str(tg) if form == '%s' and m > 1 else form % ', '.join(map(str, tg))
        if m == 1 and form == '%s':
            tgs = str(tg[0])
        return '%s %s ; %s ; %s' % ('meas' if rng.random() < 0.6 else 'peek', name, basis, tgs)
    if r < 0.78 and n >= 2:                        # disc
        m = int(rng.integers(1, n))
        tg = [int(x) for x in rng.choice(n, m, replace=False)]
        return ('disc', m, 'disc %s' % tg)
    if r < 0.90:                                   # qset of a sub-register
        if n >= 2 and rng.random() < 0.3:
            a, b = sorted(int(x) for x in rng.choice(n, 2, replace=False))     # ascending: see F13 in DESIGN.md
            return 'qset bell[%d] ; [%d, %d]' % (rng.integers(4), a, b)
        q = int(rng.integers(n))
        if rng.random() < 0.25 and n >= 2:
            a, b = rng.choice(n, 2, replace=False)
            return 'qset %s ; %s' % (ONE_Q[int(rng.integers(4))], pv(rng, [str(int(a)), str(int(b))]))
        if rng.random() < 0.2:
            return 'qset %s ; %d' % (pv(rng, [ONE_Q[0], ONE_Q[2]]), q)
        return 'qset %s ; %d' % (ONE_Q[int(rng.integers(4))], q)
    name = 'c%d' % len(names)
    names.append(name)
    return 'cdef %s ; %s' % (name, rng.choice(['np_trace(state)', 'np_real(state[0][0])', 'state.shape[0]']))


BIG_MAX = 6
BIG = False        # --big: 5-6 qubit registers; `swap` and off-slot controls stay out (the reference builds wrong unitaries for them at
                   # n >= 5), everything else (gates, slot-aligned controls, ProbVal arguments, qset / disc / meas /
                   # peek, error lines) as below


def _valid_for_big(n, ln):
    if isinstance(ln, tuple):
        return True
    if ln.startswith('swap'):
        return False
    if ln.startswith('gate '):
        parts = [x.strip() for x in ln.split(';')]
        if len(parts) >= 3:
            if 'ProbVal' in parts[2]:
                return False                     # ProbVal control lists: off-slot in general
            try:
                cs = list(eval(parts[2]))
            except Exception:                    # noqa: BLE001
                return False
            if cs:
                if 'ProbVal' in parts[1]:
                    return False
                t = int(parts[1])
                if not slot_aligned(n, t, cs):
                    return False
    return True


def program(seed, extra=False):
    rng = np.random.default_rng(50_000 + seed)
    n = int(rng.integers(5, BIG_MAX + 1)) if BIG else int(rng.integers(1, 5))
    names = []
    lines = [initial(rng, n)]
    for _ in range(int(rng.integers(2, 12))):
        ln = line(rng, n, names, extra)
        while BIG and n >= 5 and not _valid_for_big(n, ln):
            ln = line(rng, n, names, extra)
        if isinstance(ln, tuple):
            _, m, ln = ln
            n -= m
        lines.append(ln)
        if n < 1:
            break
    if n >= 1 and rng.random() < 0.25:              # one invalid line: the error text and the exit must match too
        bad = ['gate hadamardGate ; %d' % n, 'gate hadamardGate ; -1', 'gate pauliXGate ; 0 ; [%d]' % n, 'gate pauliXGate ; 0 ; [0]',
               'gate qftGate(2) ; %d' % (n - 1), 'swap 0 ; %d' % n, 'swap 0', 'disc [%d]' % n, 'disc [0, %d]' % n,
               'qset bell[0] ; [0]', 'qset comp[0] ; [%d]' % n, 'meas mm ; comp ; [%d]' % n, 'meas mm ; bell ; [0]',
               "gate hadamardGate ; 'a'", 'gate np_ones((3, 3)) ; 0', 'cdef zz ; 1/0', 'cdef zz ; undefined_name', 'frob 1',
               'gate', 'meas mm', 'jump nowhere', 'qset tensorProd(comp[0], comp[0], comp[0], comp[0], comp[0]) ; [0]',
               'gate hadamardGate ; 0 ; [] ; []', 'gate hadamardGate ; 0.5', 'swap 0 ; 0']
        if extra:
            bad.append('disc %s' % list(range(n)))      # a 0-qubit register: legal in the reference, kept out of the fixtures
        if BIG:
            bad = [b for b in bad if not b.startswith('swap')]
        # (--big: at the end only -- the templates use the FINAL register size; earlier in the program, before a `disc`, some of
        # them are valid lines with an off-slot control)
        lines.insert(len(lines) if BIG else int(rng.integers(1, len(lines) + 1)), bad[int(rng.integers(len(bad)))])
    return '\n'.join(lines), names


def run(execute, text, **kw):
    buf = io.StringIO()
    exited = False
    ns = {}
    try:
        with redirect_stdout(buf):
            ns = execute(text, **kw)
    except SystemExit:
        exited = True
    return ns, buf.getvalue(), exited


LAST_DIGIT = []      # results whose outcome weights differ from the reference's in the 15th decimal place only


def probs_agree(a, b):
    """Outcome weights are rounded to 15 decimal places by MeasurementResult (qbot/measurement.py:18-29); the
    reference gets them from abs(trace(rho_A P_i)), the backend from the diagonal of the rotated rho_A -- two
    summation orders, so about one result in a thousand lands on the other side of a rounding boundary (1e-15).
    Bit-identical lists pass silently; a difference <= 2e-15 passes and is counted; anything else is a difference
    (the contract is 1e-12 relative, north_star)."""
    a, b = [float(x) for x in a], [float(x) for x in b]
    if a == b:
        return True
    if len(a) != len(b) or max(abs(x - y) for x, y in zip(a, b)) > 2e-15:
        return False
    LAST_DIGIT.append((a, b))
    return True


_NUM = re.compile(r'-?\d+\.\d+(?:e-?\d+)?')


def stdout_agree(a, b):
    """`cout` of a measurement result prints its weights (and the same as percentages): identical text, or text that
    differs only in numbers that agree up to the 15th decimal place (2e-13 for the percentages) -- see probs_agree"""
    if a == b:
        return True
    if _NUM.sub('#', a) != _NUM.sub('#', b):
        return False
    xs, ys = _NUM.findall(a), _NUM.findall(b)
    if any(abs(float(x) - float(y)) > 2e-13 for x, y in zip(xs, ys)):
        return False
    LAST_DIGIT.append((a, b))
    return True


def compare(seed, ref_exec, our_exec, FakeState, extra=False):
    text, names = program(seed, extra)
    rns, rout, rexit = run(ref_exec, text)
    ons, oout, oexit = run(our_exec, text, state_cls=FakeState)
    if rexit != oexit or not stdout_agree(rout, oout):
        return text, 'stdout/exit differ:\n--- ref\n%s\n--- ours\n%s' % (rout, oout)
    if rexit:
        return text, None
    a, b = np.asarray(rns['state']), np.asarray(ons['state'])
    if a.shape != b.shape or not np.allclose(a, b, rtol=0, atol=1e-12):
        return text, 'state differs (max %g)' % (np.max(np.abs(a - b)) if a.shape == b.shape else -1)
    for v in names:
        x, y = rns.get(v), ons.get(v)
        if hasattr(x, 'unMeasuredDensity'):
            if list(x.basisSymbols) != list(y.basisSymbols) or not probs_agree(x.probs, y.probs):
                return text, '%s: probs / symbols differ: %s vs %s' % (v, x.probs, y.probs)
            if not np.allclose(np.asarray(x.unMeasuredDensity), np.asarray(y.unMeasuredDensity), rtol=0, atol=1e-12):
                return text, '%s: unmeasured density differs' % v
            xs, ys = x.newState, y.newState
            if (xs is None) != (ys is None) or (xs is not None and not np.allclose(np.asarray(xs), np.asarray(ys), rtol=0, atol=1e-12)):
                return text, '%s: newState differs' % v
        elif hasattr(x, 'probs') and hasattr(x, 'values'):
            if not hasattr(y, 'probs') or list(x.probs) != list(y.probs) or len(x.values) != len(y.values):
                return text, '%s: ProbVal differs' % v
            for p, q in zip(x.values, y.values):
                if hasattr(p, 'probs') and hasattr(p, 'basisSymbols'):
                    if not probs_agree(p.probs, q.probs):
                        return text, '%s: branch probs differ %s vs %s' % (v, p.probs, q.probs)
                elif not np.allclose(np.asarray(p, dtype=complex), np.asarray(q, dtype=complex), rtol=0, atol=1e-12):
                    return text, '%s: ProbVal value differs' % v
        else:
            if not np.allclose(np.asarray(x, dtype=complex), np.asarray(y, dtype=complex), rtol=0, atol=1e-12):
                return text, '%s: %r vs %r' % (v, x, y)
    return text, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--seeds', default='0:200')
    ap.add_argument('--emit')
    ap.add_argument('--show', type=int)
    ap.add_argument('--extra', action='store_true', help='also result-dependent conditions, cout of results, ProbVal arithmetic')
    ap.add_argument('--big-max', type=int, default=6, help='largest register of --big (default 6)')
    ap.add_argument('--big', action='store_true', help='5-6 qubit registers inside the validity domain of the reference at that size (no swap, slot-aligned controls only)')
    ap.add_argument('--installed', action='store_true',
                    help="second arm = the REAL reference's executeTxt with the ops installed into it (qbot_b200.install) "
                         "instead of this repo's interpreter mirror")
    a = ap.parse_args()
    global BIG, BIG_MAX
    BIG = a.big
    BIG_MAX = a.big_max
    from qbot.interpreter import executeTxt as ref_exec
    import qbot_b200
    from fake_backend import FakeState
    our_exec = qbot_b200.executeTxt
    if a.installed:
        import qbot_b200.integration as integ

        def our_exec(text, state_cls=None):
            integ.install(state_cls=state_cls)
            try:
                return ref_exec(text)
            finally:
                integ.uninstall()
    if a.show is not None:
        print(program(a.show, a.extra)[0])
        print(compare(a.show, ref_exec, our_exec, FakeState, a.extra)[1])
        return
    lo, hi = (int(x) for x in a.seeds.split(':'))
    bad = 0
    emitted = []
    for seed in range(lo, hi):
        try:
            text, why = compare(seed, ref_exec, our_exec, FakeState, a.extra)
        except Exception as e:  # noqa: BLE001
            text, why = program(seed, a.extra)[0], 'exception: %r' % e
        if why:
            bad += 1
            print('seed %d:\n%s\n=> %s\n' % (seed, text, why), flush=True)
        else:
            emitted.append(dict(name='fuzz_%d' % seed, text=text, vars=program(seed, a.extra)[1]))
    print('seeds %d:%d: %d differences (%d results differ in the 15th decimal place of a weight only)' % (lo, hi, bad, len(LAST_DIGIT)))
    if a.emit:
        with open(a.emit, 'w') as f:
            json.dump(emitted, f, indent=0)
    sys.exit(1 if bad else 0)


if __name__ == '__main__':
    main()
