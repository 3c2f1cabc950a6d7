"""CPU fuzzer of the planner + sweep specialiser (no GPU): random gate lists -> qt_plan_best ->
qj_generate -> g++ -> execution, compared with the oracle's definitional gate application.

    python scripts/fuzz_codegen.py --seeds 0:200 --jobs 8

Each case draws a generator (rc circuits, mixed tileable gates, flip-heavy traffic, CX-then-CH), a
register size close to the tile size (so controls are thread / tile bits), the tile size M, the
register bits per stage R, and the plan-search effort.  Failures are printed with everything needed
to reproduce them (`--one SEED`).  TEST INFRASTRUCTURE: uses oracle/ as the checker."""
import argparse
import os
import sys
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def case(seed):
    import numpy as np
    rng = np.random.default_rng(10_000 + seed)
    R = int(rng.choice([4, 5]))
    M = int(rng.choice([11, 12]))
    trials = int(rng.choice([1, 8, 32, 128]))
    kind = int(rng.integers(0, 5))
    n = int(rng.integers(M, M + 5))
    os.environ['QBOT_B200_PLAN_R'] = str(R)
    os.environ['QBOT_B200_PLAN_TRIALS'] = str(trials)
    import jit_emu
    import plan_emu
    from qbot_b200.circuits import rc
    from test_planner import random_gate_list, oracle_apply_bits, rand_ket
    from test_jit_codegen import tileable, _flip_heavy, _cx_then_ch
    if kind == 0:
        gl = plan_emu.circuit_to_bits(n, rc(n, int(rng.integers(3, 14)), seed))
    elif kind == 1:
        gl = tileable(random_gate_list(rng, n, int(rng.integers(20, 90))))
    elif kind == 2:
        gl = _flip_heavy(rng, n, int(rng.integers(20, 90)))
    elif kind == 3:
        gl = _cx_then_ch(rng, n, int(rng.integers(5, 30)))
    else:
        gl = plan_emu.circuit_to_bits(n, rc(n, int(rng.integers(2, 8)), seed)) + _flip_heavy(rng, n, 30) + \
            tileable(random_gate_list(rng, n, 20))
    psi = rand_ket(rng, n)
    desc = dict(seed=seed, kind=kind, n=n, M=M, R=R, trials=trials, gates=len(gl))
    try:
        out, info = jit_emu.run(n, gl, psi, M=M)
    except AssertionError as e:
        if 'does not run' in str(e):
            return ('skip', desc, str(e))
        return ('error', desc, repr(e))
    except Exception as e:  # noqa: BLE001
        return ('error', desc, repr(e))
    ref = psi
    for m, tb, cm in gl:
        ref = oracle_apply_bits(ref, n, m, tb, cm)
    err = float(np.max(np.abs(out - ref)))
    desc.update(info)
    return ('ok' if err < 1e-12 else 'MISMATCH', desc, err)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--seeds', default='0:64')
    ap.add_argument('--jobs', type=int, default=os.cpu_count())
    ap.add_argument('--one', type=int)
    a = ap.parse_args()
    if a.one is not None:
        print(case(a.one))
        return
    lo, hi = (int(x) for x in a.seeds.split(':'))
    import plan_emu
    plan_emu.lib()                                        # build once, before the workers race for it
    bad = 0
    with ProcessPoolExecutor(a.jobs) as ex:
        for status, desc, extra in ex.map(case, range(lo, hi)):
            if status != 'ok':
                bad += status != 'skip'
                print(status, desc, extra, flush=True)
    print(f"seeds {lo}:{hi}: {bad} failures")
    sys.exit(1 if bad else 0)


if __name__ == '__main__':
    main()
