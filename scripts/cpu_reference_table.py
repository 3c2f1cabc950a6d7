"""The CPU numbers SURVEY.md 8(d) asks for next to the GPU ones ("How the CPU reference is timed"), host cores only:

  (i)   the STOCK reference's `gate` op, line by line through its interpreter (baseline/_ref, or /root/reference in the
        build container), on rc(n_ref, D, seed = n_ref) for n_ref in {8, 10, 12} -- the largest registers its 4^n
        representation allows -- with all BLAS threads and with one;
  (ii)  "reference arithmetic on a ket": genGateForFullHilbertSpace(n, t, G) @ psi (qbot/qgates.py:161-182) for
        n in {10, 12, 13} -- NOT a path the reference has (it keeps density matrices only), labelled as such;
  (iii) the extrapolation of (i) to the benchmark sizes (x8 flops and x4 bytes per qubit for rho), labelled as such: the
        reference cannot run there.

    python scripts/cpu_reference_table.py [--budget 40] [--json profiles/r02_cpu_reference_table.json]

One warm-up gate per case, then gates in circuit order until the budget is spent (n = 8, 10: median of 5 passes over the
same gates).  Lines the stock op cannot execute (TypeError for some multi-control layouts) are skipped and
counted.  TEST / MEASUREMENT INFRASTRUCTURE: nothing here is on the product path."""
import argparse
import io
import json
import os
import statistics
import subprocess
import sys
import time
from contextlib import redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def stock_gate_op(n, depth, budget_s, passes):
    import numpy as np
    import bench_configs as bc
    from qbot_b200 import circuits
    ref = bc.load_reference()
    if ref is None:
        return None
    interp = ref['interp']
    gates = circuits.rc(n, depth, n)
    lines = [f"qset tensorExp(comp[0], {n})"] + [g.dsl() for g in gates]
    per_pass = []
    done = skipped = 0
    for _ in range(passes):
        ns = {'state': np.array([], dtype=complex), '__updated_state': False, '__marks': dict(), '__prev_jump': -1}
        interp.runtime(ns, lines, 0, 1)
        interp.runtime(ns, lines, 1, 2)                 # warm-up gate
        t_total, done, skipped = 0.0, 0, 0
        for i in range(2, len(lines)):
            t0 = time.perf_counter()
            try:
                with redirect_stdout(io.StringIO()):
                    interp.runtime(ns, lines, i, i + 1)
                t_total += time.perf_counter() - t0
                done += 1
            except SystemExit:
                skipped += 1
            if t_total > budget_s:
                break
        per_pass.append(done / max(t_total, 1e-9))
    kinds = {}
    for g in gates[1:1 + done + skipped]:
        k = {0: 'uncontrolled', 1: 'cnot', 2: 'toffoli'}[len(g.controls)]
        kinds[k] = kinds.get(k, 0) + 1
    return dict(qubits=n, gates_timed=done, skipped_F6=skipped, passes=passes, gates_per_s=statistics.median(per_pass),
                seconds_per_gate=1 / statistics.median(per_pass), gate_mix=kinds,
                what=f"stock reference `gate` op on rc({n}, {depth}, {n}), {1 << n} x {1 << n} complex128 density matrix")


def ket_arithmetic(n, budget_s):
    import numpy as np
    import bench_configs as bc
    from qbot_b200 import circuits
    ref = bc.load_reference()
    if ref is None:
        return None
    import importlib
    rg = importlib.import_module('qbot.qgates')
    psi = np.zeros(1 << n, dtype=complex)
    psi[0] = 1
    gates = [g for g in circuits.rc(n, 4, n) if not g.controls]
    t_total, done = 0.0, 0
    psi = rg.genGateForFullHilbertSpace(n, gates[0].target, gates[0].matrix()) @ psi
    for g in gates[1:]:
        t0 = time.perf_counter()
        psi = rg.genGateForFullHilbertSpace(n, g.target, g.matrix()) @ psi
        t_total += time.perf_counter() - t0
        done += 1
        if t_total > budget_s:
            break
    return dict(qubits=n, gates_timed=done, gates_per_s=done / t_total, seconds_per_gate=t_total / done,
                what=f"genGateForFullHilbertSpace({n}, t, G) @ psi, uncontrolled gates of rc({n}, 4, {n}) -- reference arithmetic on a "
                     f"ket, not a path the reference has (SURVEY F1)")


def one_run(budget):
    import bench_configs as bc
    used = bc.blas_threads(int(os.environ.get('QB_TABLE_THREADS', os.cpu_count())))
    out = dict(blas_threads=used, host_cores=os.cpu_count(), stock=[], ket=[])
    for n, depth, passes in ((8, 4, 5), (10, 3, 5), (12, 50, 1)):
        r = stock_gate_op(n, depth, budget if n == 12 else min(budget, 10.0), passes)
        if r:
            out['stock'].append(r)
    for n in (10, 12, 13):
        r = ket_arithmetic(n, min(budget, 15.0))
        if r:
            out['ket'].append(r)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--budget', type=float, default=40.0)
    ap.add_argument('--json')
    ap.add_argument('--child', action='store_true')
    a = ap.parse_args()
    if a.child:
        print(json.dumps(one_run(a.budget)))
        return
    runs = []
    for threads in (os.cpu_count(), 1):
        env = dict(os.environ, QB_TABLE_THREADS=str(threads), OPENBLAS_NUM_THREADS=str(threads), OMP_NUM_THREADS=str(threads))
        p = subprocess.run([sys.executable, os.path.abspath(__file__), '--child', '--budget', str(a.budget)], env=env,
                           capture_output=True, text=True, check=True)
        runs.append(json.loads(p.stdout.strip().splitlines()[-1]))
    # (iii) extrapolation from the 12-qubit figure with all threads: x8 flops per qubit (two 2^n-cubed zgemms)
    s12 = next((r for r in runs[0]['stock'] if r['qubits'] == 12), None)
    extra = None
    if s12:
        extra = {f"{n} qubits": {"seconds_per_gate": s12['seconds_per_gate'] * 8.0 ** (n - 12),
                                 "rho_bytes": 16 * 4 ** n} for n in (20, 30, 34)}
    table = dict(runs=runs, extrapolation_from_12_qubits=extra,
                 note="EXTRAPOLATION (x8 flops, x4 bytes per qubit for rho): the reference cannot run at these sizes -- a 20-qubit "
                      "density matrix alone is 16 TiB")
    txt = json.dumps(table, indent=1)
    if a.json:
        with open(a.json, 'w') as f:
            f.write(txt + "\n")
    for run in runs:
        print(f"== BLAS threads {run['blas_threads']} of {run['host_cores']} host cores")
        for r in run['stock']:
            print(f"  stock gate op   n={r['qubits']:2d}: {r['gates_per_s']:10.3f} gates/s  ({r['seconds_per_gate'] * 1e3:10.2f} ms per gate; "
                  f"{r['gates_timed']} gates, {r['skipped_F6']} skipped, mix {r['gate_mix']})")
        for r in run['ket']:
            print(f"  ket arithmetic  n={r['qubits']:2d}: {r['gates_per_s']:10.3f} gates/s  ({r['seconds_per_gate'] * 1e3:10.2f} ms per gate; {r['gates_timed']} gates)")
    if extra:
        for k, v in extra.items():
            print(f"  EXTRAPOLATED    {k}: {v['seconds_per_gate']:.3g} s per gate, rho = {v['rho_bytes'] / 2 ** 40:.3g} TiB (cannot exist)")


if __name__ == '__main__':
    main()
