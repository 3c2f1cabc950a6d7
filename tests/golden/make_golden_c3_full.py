#!/usr/bin/env python3
"""Golden fixture for BASELINE config 3 at its FULL size, from the REAL reference (build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_c3_full.py        # ~25 min on 8 cores

    c3_12.npz / c3_12.json
      the 12-qubit density-matrix program of config 3 (rc(12, 50, 12), a `meas` of two random qubits and a
      ProbVal-target Hadamard after every 10 layers, final `disc` of 4 qubits) as a loop over the
      reference's own state functions on its 4096 x 4096 matrix -- the same loop as `c3_6` / `c3_8` in
      make_golden_configs.py, at the size `bench.py`'s `configs.c3` sub-line and `--impl reference` run.
      Stored: the final 8-qubit register (256 x 256), every measurement's probabilities, and of the
      12-qubit register just before the `disc`: its diagonal, 2 sampled rows, trace and purity.
(Reproducibility: the 4096-cubed zgemms run on all BLAS threads, so a re-run on another machine / thread count reproduces the
arrays to rounding (~1e-15 of the largest entry), not bit for bit; the recorded probabilities are rounded to 15 decimals by the
reference itself.  Recorded here with numpy 2.3 / OpenBLAS on 8 threads.)
Unitaries come from the reference's builders inside their validity domain and from the definitional
unitary outside it, exactly as in make_golden_configs.py.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, '/root/reference')
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import qbot.qgates as rg                      # noqa: E402  (the reference)
import qbot.density as rd                     # noqa: E402
import qbot.measurement as rm                 # noqa: E402
import qbot.basis as rb                       # noqa: E402

from qbot_b200 import circuits                # noqa: E402
from make_golden_configs import unitary       # noqa: E402

ROWS = [0, 2730]


def main():
    n, depth, seed = 12, 50, 12
    H = circuits.HADAMARD
    ops = circuits.c3_ops(n, depth, seed)
    rho = np.zeros((1 << n, 1 << n), dtype=complex)
    rho[0, 0] = 1
    probs, arr = {}, {}
    built = 0
    t0 = time.perf_counter()
    for i, op in enumerate(ops):
        if op.kind == 'gate':
            u, b = unitary(n, op.gate)
            built += b
            rho = rg.applyGate(u, rho)
        elif op.kind == 'meas':
            res = rm.measureArbitraryMultiState(rho, rb.computation, list(op.qubits), True)
            probs[op.name] = [float(p) for p in res.probs]
            rho = res.newState
        elif op.kind == 'pgate':        # operators.py:308-316: ProbVal of unitaries -> per-branch applyGate -> ensemble
            branches = [rg.applyGate(rg.genGateForFullHilbertSpace(n, t, H), rho) for t in op.qubits]
            rho = rd.densityEnsambleToDensity([.5, .5], branches)
        else:                           # disc keeps the qubits that are NOT listed (operators.py:169-175)
            arr['c3_12_before_disc_diag'] = np.diag(rho).copy()
            arr['c3_12_before_disc_rows'] = rho[ROWS].copy()
            tr, pur = np.trace(rho), np.vdot(rho.conj().T, rho)       # tr(rho^2) without a 4096^3 product
            _, rho = rd.partialTraceArbitrary(rho, n, list(op.qubits))
        if i % 10 == 0:
            print(f"op {i}/{len(ops)} {op.kind} {time.perf_counter() - t0:.0f}s", flush=True)
    arr['c3_12_state'] = rho
    meta = dict(n=n, depth=depth, seed=seed, ops=len(ops), probs=probs, reference_built_unitaries=built, rows=ROWS,
                before_disc_trace=[float(tr.real), float(tr.imag)], before_disc_purity=float(pur.real),
                seconds=round(time.perf_counter() - t0))
    np.savez_compressed(os.path.join(HERE, 'c3_12.npz'), **arr)
    with open(os.path.join(HERE, 'c3_12.json'), 'w') as f:
        json.dump({'c3_12': meta}, f, indent=0)
    print('written', {k: v.shape for k, v in arr.items()}, meta['seconds'], 's')


if __name__ == '__main__':
    main()
