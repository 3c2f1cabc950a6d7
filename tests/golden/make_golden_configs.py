#!/usr/bin/env python3
"""Golden fixtures for the BASELINE configs at the sizes SURVEY.md 8(d) names, from the REAL reference
(build container only; the reference does not travel to the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_configs.py

    configs.npz / configs.json
      rc_10_200_20    rho after rc(10, 200, 20) pushed through the reference's applyGate: its diagonal,
                      8 sampled rows, trace and purity (the full 1024 x 1024 matrix would be 16 MiB)
      c3_6, c3_8      config 3 at n = 6 / 8: mid-circuit `meas` (reference collapse), one ProbVal-target
                      Hadamard, final `disc`, as a loop over the reference's own state functions
                      (applyGate, measureArbitraryMultiState, densityEnsambleToDensity,
                      partialTraceArbitrary) -- final register and every measurement's probabilities
      c4_64_6         config 4 at B = 64, n = 6: per branch the reference's applyGate + measurement ->
                      [64, 16] outcome weights
Unitaries come from the reference's builders inside their validity domain and from the definitional
unitary outside it -- the `gate` op of the stock DSL is wrong there at n >= 5, which
is why these are loops over the state functions and not `executeTxt` runs.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, '/root/reference')
sys.path.insert(0, ROOT)

import qbot.qgates as rg                      # noqa: E402  (the reference)
import qbot.density as rd                     # noqa: E402
import qbot.measurement as rm                 # noqa: E402
import qbot.basis as rb                       # noqa: E402

from qbot_b200 import circuits                # noqa: E402
from oracle import qbot_oracle as orc         # noqa: E402  (definitional U where the reference is invalid)
from make_golden import slot_aligned          # noqa: E402


def unitary(n, g):
    built = 0
    if len(g.controls) == 0:
        u, built = rg.genGateForFullHilbertSpace(n, g.target, g.matrix()), 1
    elif slot_aligned(n, g.target, list(g.controls)):
        u, built = rg.genMultiControlledGate(n, list(g.controls), g.target, g.matrix()), 1
    else:
        u = orc.controlled_unitary(n, g.controls, g.target, g.matrix())
    return u, built


def main():
    arr, meta = {}, {}
    # ---- config 2 parity run: rc(10, 200, 20) ------------------------------------------------
    n = 10
    gates = circuits.rc(n, 200, 20)
    rho = np.zeros((1 << n, 1 << n), dtype=complex)
    rho[0, 0] = 1
    nref = 0
    for g in gates:
        u, b = unitary(n, g)
        nref += b
        rho = rg.applyGate(u, rho)
    rows = [0, 1, 2, 341, 512, 682, 1000, 1023]
    arr['rc_10_200_20_diag'] = np.diag(rho).copy()
    arr['rc_10_200_20_rows'] = rho[rows].copy()
    meta['rc_10_200_20'] = dict(n=n, depth=200, seed=20, gates=len(gates), reference_built_unitaries=nref, rows=rows,
                                trace=[float(np.trace(rho).real), float(np.trace(rho).imag)],
                                purity=float(np.trace(rho @ rho).real))
    # ---- config 3 at n = 6, 8 ---------------------------------------------------------------------
    H = circuits.HADAMARD
    for n, depth in ((6, 20), (8, 30)):
        ops = circuits.c3_ops(n, depth, n)
        rho = np.zeros((1 << n, 1 << n), dtype=complex)
        rho[0, 0] = 1
        probs = {}
        for op in ops:
            if op.kind == 'gate':
                rho = rg.applyGate(unitary(n, op.gate)[0], rho)
            elif op.kind == 'meas':
                res = rm.measureArbitraryMultiState(rho, rb.computation, list(op.qubits), True)
                probs[op.name] = [float(p) for p in res.probs]
                rho = res.newState
            elif op.kind == 'pgate':        # operators.py:308-316: ProbVal of unitaries -> per-branch applyGate -> ensemble
                branches = [rg.applyGate(rg.genGateForFullHilbertSpace(n, t, H), rho) for t in op.qubits]
                rho = rd.densityEnsambleToDensity([.5, .5], branches)
            else:                           # disc keeps the qubits that are NOT listed (operators.py:169-175)
                _, rho = rd.partialTraceArbitrary(rho, n, list(op.qubits))
        arr[f'c3_{n}_state'] = rho
        meta[f'c3_{n}'] = dict(n=n, depth=depth, seed=n, probs=probs, ops=len(ops))
    # ---- config 4 at B = 64, n = 6 ------------------------------------------------------------------
    B, n = 64, 6
    factors, w, gates, ang, tgt, measured = circuits.c4_inputs(B, n, 16)
    out = np.zeros((B, 1 << len(measured)))
    us = [unitary(n, g)[0] for g in gates]
    for b in range(B):
        psi = np.array([1.0 + 0j])
        for q in range(n):
            psi = np.kron(psi, factors[b, q])
        rho = np.outer(psi, psi.conj())
        for u in us:
            rho = rg.applyGate(u, rho)
        rho = rg.applyGate(rg.genGateForFullHilbertSpace(n, int(tgt[b]), circuits.z_rot(float(ang[b]))), rho)
        res = rm.measureArbitraryMultiState(rho, rb.computation, list(measured), False)
        out[b] = res.probs
    arr['c4_64_6_probs'] = out
    meta['c4_64_6'] = dict(B=B, n=n, seed=16, measured=measured)
    np.savez_compressed(os.path.join(HERE, 'configs.npz'), **arr)
    with open(os.path.join(HERE, 'configs.json'), 'w') as f:
        json.dump(meta, f, indent=0)
    print('written', {k: v.shape for k, v in arr.items()})


if __name__ == '__main__':
    main()
