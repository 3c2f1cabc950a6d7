"""Differential fuzzer of the KET-MODE register against the density-matrix register (no GPU): the random programs of
scripts/fuzz_dsl.py (every state op with plain and ProbVal arguments, conditions, sub-register qset, disc, meas / peek in three
bases, injected error lines; n <= 4) run twice through this repo's ops on the numpy double --

  (a) as generated: the register starts as a product of DENSITY matrices (`comp[0]`, `bell[2]`, ...), the reference's only
      representation -- the arm scripts/fuzz_dsl.py compares with the live reference;
  (b) with the first line's factors replaced by the corresponding KETS (`comp.kets[0]`, `bell.kets[2]`, ...): `qset` of a 1-D
      array gives a ket-mode register (the new representation), which every op has to treat as psi psi^dagger --
      in place where it can (gates, swap, peek), by converting where the result is mixed (meas, ProbVal gates, sub-register
      qset, disc).

stdout, exit behaviour, the final register (as a density matrix) and every named result must agree.

    python scripts/fuzz_ket_vs_dm.py --seeds 0:2000 [--extra]

TEST INFRASTRUCTURE."""
import argparse
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'scripts'))

_FACTOR = re.compile(r'\b(comp|hada|bell)\[(\d)\]')


def as_density(a):
    a = np.asarray(a)
    return np.outer(a, a.conj()) if a.ndim == 1 and a.size else a


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--seeds', default='0:500')
    ap.add_argument('--extra', action='store_true')
    a = ap.parse_args()
    import fuzz_dsl as fd
    import qbot_b200
    from fake_backend import FakeState
    lo, hi = (int(x) for x in a.seeds.split(':'))
    bad = kets = 0
    for seed in range(lo, hi):
        text, names = fd.program(seed, a.extra)
        # (expressions that read the register by name see a 1-D ket in arm (b): inherent to the representation, SURVEY.md F11 -- left out)
        lines = [ln for ln in text.split('\n') if not (ln.startswith('cdef') and 'state' in ln)]
        text = '\n'.join(lines)
        ktext = '\n'.join([_FACTOR.sub(r'\1.kets[\2]', lines[0])] + lines[1:])
        dns, dout, dexit = fd.run(qbot_b200.executeTxt, text, state_cls=FakeState)
        kns, kout, kexit = fd.run(qbot_b200.executeTxt, ktext, state_cls=FakeState)
        why = None
        # error windows quote the program text: compare them with the first line masked
        mask = lambda s: _FACTOR.sub(r'\1[\2]', s.replace('.kets[', '['))      # noqa: E731
        if dexit != kexit or not fd.stdout_agree(mask(dout), mask(kout)):
            why = 'stdout / exit differ:\n--- density\n%s\n--- ket\n%s' % (dout, kout)
        elif not dexit:
            x, y = as_density(dns['state']), as_density(kns['state'])
            kets += np.asarray(kns['state']).ndim == 1
            if x.shape != y.shape or not np.allclose(x, y, rtol=0, atol=1e-12):
                why = 'register differs (max %g)' % (np.max(np.abs(x - y)) if x.shape == y.shape else -1)
            for v in names if why is None else []:
                p, q = dns.get(v), kns.get(v)
                if p is None and q is None:          # (a left-out `cdef`)
                    continue
                if hasattr(p, 'unMeasuredDensity'):
                    if list(p.basisSymbols) != list(q.basisSymbols) or not fd.probs_agree(p.probs, q.probs):
                        why = '%s: probs / symbols differ %s vs %s' % (v, p.probs, q.probs)
                    elif not np.allclose(np.asarray(p.unMeasuredDensity), np.asarray(q.unMeasuredDensity), rtol=0, atol=1e-12):
                        why = '%s: unMeasuredDensity differs' % v
                    elif (p.newState is None) != (q.newState is None) or (p.newState is not None and not np.allclose(
                            as_density(p.newState), as_density(q.newState), rtol=0, atol=1e-12)):
                        why = '%s: newState differs' % v
                elif hasattr(p, 'probs') and hasattr(p, 'values'):
                    if not hasattr(q, 'probs') or list(p.probs) != list(q.probs):
                        why = '%s: ProbVal differs' % v
                else:
                    try:
                        same = np.allclose(np.asarray(p, dtype=complex), np.asarray(q, dtype=complex), rtol=0, atol=1e-12)
                    except Exception:      # noqa: BLE001
                        same = False
                    if not same:
                        why = '%s: %r vs %r' % (v, p, q)
                if why:
                    break
        if why:
            bad += 1
            print('seed %d:\n%s\n=> %s\n' % (seed, ktext, why), flush=True)
    print('seeds %d:%d: %d differences (%d programs ended on a ket-mode register)' % (lo, hi, bad, kets))
    sys.exit(1 if bad else 0)


if __name__ == '__main__':
    main()
