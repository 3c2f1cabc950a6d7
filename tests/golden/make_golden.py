#!/usr/bin/env python3
"""Generate the golden fixtures in this directory from the REAL reference.

Run in the build container only (the reference lives at /root/reference there and does not
travel to the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Outputs (small, committed):
    l1_cases.npz / l1_cases.json   -- qgates / density / measurement functions on seeded inputs
    scripts.json + scripts.npz     -- DSL programs run through the reference's executeTxt
    scripts_fuzz.json + .npz       -- random DSL programs (scripts/fuzz_dsl.py --emit ...scripts_fuzz_src.json), same recording
    scripts_fuzz_big.json + .npz   -- the same on 5-6 qubit registers (scripts/fuzz_dsl.py --big --emit ...scripts_fuzz_big_src.json)
    probval.json                   -- ProbVal normalise / funcWrapper ordering cases
    rc_small.npz                   -- rc(n, D, seed) circuits pushed through the reference's applyGate

Nothing here is product code; the product and the oracle never read /root/reference.
"""
import io
import json
import os
import sys
from contextlib import redirect_stdout

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
from qbot.probVal import ProbVal, funcWrapper  # noqa: E402
from qbot.interpreter import executeTxt       # noqa: E402
from qbot.evaluation import globalNameSpace   # noqa: E402

from qbot_b200.circuits import rc             # noqa: E402
from oracle import qbot_oracle as orc         # noqa: E402  (only for the definitional U where the reference is invalid)


def rand_density(rng, n, rank=3):
    dim = 1 << n
    rho = np.zeros((dim, dim), dtype=complex)
    w = rng.random(rank) + 0.1
    w /= w.sum()
    for p in w:
        v = rng.normal(size=dim) + 1j * rng.normal(size=dim)
        v /= np.linalg.norm(v)
        rho += p * np.outer(v, v.conj())
    return rho


def rand_unitary(rng, k):
    dim = 1 << k
    a = rng.normal(size=(dim, dim)) + 1j * rng.normal(size=(dim, dim))
    q, r = np.linalg.qr(a)
    return q * (np.diag(r) / np.abs(np.diag(r)))


BASES = {'comp': rb.computation, 'hada': rb.hadamard, 'bell': rb.bell}


def slot_aligned(n, t, controls):
    """Layouts where the reference's genMultiControlledGate is known to be right (F6)."""
    c = len(controls)
    if c == 1 and n <= 4:
        return True
    if t == 0:
        return list(controls) == list(range(n - c, n))
    return list(controls) == list(range(t - c, t))


def record_scripts(src_name, out_stem):
    """run every program of `src_name` through the reference's executeTxt and record stdout, exit
    behaviour, the final register and the named results -> <out_stem>.json / .npz"""
    scripts = json.load(open(os.path.join(HERE, src_name)))
    sarrays = {}
    sout = []
    for i, sc in enumerate(scripts):
        buf = io.StringIO()
        exited = False
        try:
            with redirect_stdout(buf):
                ns = executeTxt(sc['text'])
        except SystemExit:
            exited = True
            ns = {}
        rec = dict(name=sc['name'], text=sc['text'], stdout=buf.getvalue(), exited=exited, vars={})
        if 'state' in ns and isinstance(ns['state'], np.ndarray):
            sarrays[f's{i}'] = ns['state']
            rec['state'] = f's{i}'
        for v in ([] if exited else sc.get('vars', [])):
            val = ns.get(v)
            if isinstance(val, rm.MeasurementResult):
                rec['vars'][v] = dict(type='meas', probs=[float(p) for p in val.probs], symbols=val.basisSymbols)
                sarrays[f's{i}_{v}_un'] = val.unMeasuredDensity
            elif isinstance(val, ProbVal):
                rec['vars'][v] = dict(type='probval', probs=val.probs, values=json.loads(json.dumps(val.values, default=str)))
            elif isinstance(val, np.ndarray):
                sarrays[f's{i}_{v}'] = val
                rec['vars'][v] = dict(type='array', key=f's{i}_{v}')
            elif isinstance(val, (complex, np.complexfloating)):
                sarrays[f's{i}_{v}'] = np.asarray(val)
                rec['vars'][v] = dict(type='array', key=f's{i}_{v}')
            else:
                rec['vars'][v] = dict(type='py', value=json.loads(json.dumps(val, default=str)))
        sout.append(rec)
    np.savez_compressed(os.path.join(HERE, out_stem + '.npz'), **sarrays)
    with open(os.path.join(HERE, out_stem + '.json'), 'w') as f:
        json.dump(sout, f, indent=0)
    return sout


def main():
    rng = np.random.default_rng(20261018)
    arrays = {}
    meta = []

    def put(name, a):
        arrays[name] = np.asarray(a)
        return name

    # ---- gate application (qgates.applyGate with reference-built unitaries) -------------
    cid = 0
    for n in (1, 2, 3, 4, 5):
        for k in (1, 2, 3):
            if k > n:
                continue
            for t in range(0, n - k + 1):
                g = rand_unitary(rng, k)
                rho = rand_density(rng, n)
                u = rg.genGateForFullHilbertSpace(n, t, g)
                out = rg.applyGate(u, rho)
                meta.append(dict(kind='gate', id=cid, n=n, t=t, k=k, controls=[],
                                 g=put(f'g{cid}', g), rho=put(f'i{cid}', rho), out=put(f'o{cid}', out),
                                 unitary_from='reference'))
                cid += 1
    # controlled gates: reference builder inside its validity domain, definitional outside
    layouts = [(2, 0, [1]), (2, 1, [0]), (3, 2, [0]), (3, 0, [2]), (3, 1, [0]), (3, 1, [2]),
               (3, 2, [0, 1]), (3, 0, [1, 2]), (4, 3, [1, 2]), (4, 2, [0, 1]), (4, 0, [2, 3]),
               (4, 1, [3]), (4, 3, [0]), (5, 2, [0, 1]), (5, 0, [3, 4]), (5, 4, [1, 2, 3]),
               # outside the reference's validity domain -> definitional unitary on both sides
               (3, 1, [0, 2]), (4, 0, [3, 1]), (5, 2, [4]), (5, 1, [3, 0]), (5, 4, [0])]
    for (n, t, controls) in layouts:
        g = rand_unitary(rng, 1)
        rho = rand_density(rng, n)
        if slot_aligned(n, t, controls):
            u = rg.genMultiControlledGate(n, list(controls), t, g)
            src = 'reference'
            assert np.allclose(u, orc.controlled_unitary(n, controls, t, g)), (n, t, controls)
        else:
            u = orc.controlled_unitary(n, controls, t, g)
            src = 'definitional (reference builder invalid here, SURVEY F5/F6)'
        out = rg.applyGate(u, rho)
        meta.append(dict(kind='gate', id=cid, n=n, t=t, k=1, controls=list(controls),
                         g=put(f'g{cid}', g), rho=put(f'i{cid}', rho), out=put(f'o{cid}', out),
                         unitary_from=src))
        cid += 1
    # swaps (reference genSwapGate is right only for n <= 4)
    for n in (2, 3, 4):
        for a in range(n):
            for b in range(n):
                rho = rand_density(rng, n)
                u = rg.genSwapGate(n, a, b)
                assert np.allclose(u, orc.swap_unitary(n, a, b))
                out = rg.applyGate(u, rho)
                meta.append(dict(kind='swap', id=cid, n=n, a=a, b=b, rho=put(f'i{cid}', rho),
                                 out=put(f'o{cid}', out), unitary_from='reference'))
                cid += 1
    for (n, a, b) in ((5, 0, 4), (5, 1, 3), (6, 2, 5)):
        rho = rand_density(rng, n)
        out = rg.applyGate(orc.swap_unitary(n, a, b), rho)
        meta.append(dict(kind='swap', id=cid, n=n, a=a, b=b, rho=put(f'i{cid}', rho),
                         out=put(f'o{cid}', out),
                         unitary_from='definitional (reference genSwapGate wrong for n>=5, SURVEY F5)'))
        cid += 1
    # shift gates (correct at all n)
    for n in (2, 3, 4, 5):
        for up in (True, False):
            for s in (1, 2):
                if s >= n:
                    continue
                u = rg.genShiftGate(n, up, s)
                meta.append(dict(kind='shift', id=cid, n=n, up=up, shifts=s, u=put(f'u{cid}', u.real.astype(np.int8))))
                cid += 1

    # ---- partial trace / interweave / replace -------------------------------------------
    for n, lists in ((2, [[0], [1]]), (3, [[0], [1], [2], [0, 1], [0, 2], [1, 2], [2, 0]]),
                     (4, [[1], [3], [0, 3], [1, 2], [0, 1, 3], [3, 1]]),
                     (5, [[2], [0, 4], [1, 2, 3], [4, 0, 2]])):
        for qs in lists:
            rho = rand_density(rng, n)
            a, b = rd.partialTraceArbitrary(rho, n, list(qs))
            meta.append(dict(kind='ptrace', id=cid, n=n, qubits=list(qs), rho=put(f'i{cid}', rho),
                             a=put(f'a{cid}', a), b=put(f'b{cid}', b)))
            cid += 1
    for (na, nb, pos) in ((1, 1, [0]), (1, 1, [1]), (1, 2, [1]), (2, 1, [0, 2]), (2, 2, [1, 3]),
                          (1, 3, [3]), (2, 3, [0, 4]), (3, 2, [1, 2, 4])):
        ra, rbm = rand_density(rng, na), rand_density(rng, nb)
        out = rd.interweaveDensities(ra, rbm, list(pos))
        meta.append(dict(kind='interweave', id=cid, pos=list(pos), a=put(f'a{cid}', ra), b=put(f'b{cid}', rbm),
                         out=put(f'o{cid}', out)))
        cid += 1
    for (n, k, tg) in ((2, 1, [0]), (2, 1, [1]), (3, 1, [1]), (3, 2, [0, 2]), (4, 2, [1, 2]), (4, 1, [3]),
                       (5, 2, [0, 4]), (2, 2, [0, 1])):
        rho, new = rand_density(rng, n), rand_density(rng, k)
        out = rd.replaceArbitrary(rho, new, list(tg))
        meta.append(dict(kind='replace', id=cid, n=n, targets=list(tg), rho=put(f'i{cid}', rho),
                         new=put(f'n{cid}', new), out=put(f'o{cid}', out)))
        cid += 1

    # ---- measurement ------------------------------------------------------------------
    mcases = [(1, 'comp', None), (1, 'hada', [0]), (2, 'comp', None), (2, 'bell', None), (2, 'hada', [1]),
              (2, 'bell', [1, 0]), (3, 'comp', [1]), (3, 'comp', [0, 2]), (3, 'hada', [2, 0]), (3, 'bell', [0, 1]),
              (3, 'bell', [0, 2]), (4, 'comp', [0, 3]), (4, 'bell', [1, 2]), (4, 'hada', [0, 1, 2, 3]),
              (4, 'bell', None), (5, 'comp', [4, 2]), (5, 'hada', [0]), (5, 'bell', [0, 1, 3, 4]), (4, 'comp', {0, 3})]
    for (n, bname, tg) in mcases:
        for ret in (True, False):
            rho = rand_density(rng, n)
            res = rm.measureArbitraryMultiState(rho, BASES[bname], tg, ret)
            m = dict(kind='measure', id=cid, n=n, basis=bname, targets=None if tg is None else sorted(tg),
                     targets_as_given=None if tg is None else list(tg), targets_is_set=isinstance(tg, set),
                     return_state=ret, rho=put(f'i{cid}', rho), probs=[float(p) for p in res.probs],
                     symbols=list(res.basisSymbols), unmeasured=put(f'a{cid}', res.unMeasuredDensity))
            if ret:
                m['new_state'] = put(f'o{cid}', res.newState)
            meta.append(m)
            cid += 1

    # ---- ensemble ------------------------------------------------------------------------
    for n in (1, 2, 3):
        ps = rng.random(3)
        ps /= ps.sum()
        rhos = [rand_density(rng, n) for _ in range(3)]
        out = rd.densityEnsambleToDensity(list(ps), rhos)
        meta.append(dict(kind='ensemble', id=cid, probs=[float(p) for p in ps],
                         rhos=[put(f'r{cid}_{j}', r) for j, r in enumerate(rhos)], out=put(f'o{cid}', out)))
        cid += 1

    np.savez_compressed(os.path.join(HERE, 'l1_cases.npz'), **arrays)
    with open(os.path.join(HERE, 'l1_cases.json'), 'w') as f:
        json.dump(meta, f, indent=0)

    # ---- ProbVal rules ---------------------------------------------------------------
    pv_cases = []
    for probs, vals in (([.5, .25, .25], [1, 1, 2]), ([.5, .5], [3, 4]), ([.999999, .000001], [1, 2]),
                        ([.2, .3, .5], ['a', 'b', 'a']), ([.25, .25, .25, .25], [0, 1, 0, 1]),
                        ([1 / 3, 1 / 3, 1 / 3], [1.0, 1.000001, 2.0]), ([.1, .2, .7], [(1, 2), (1, 2), (2, 1)])):
        pv = ProbVal(list(probs), list(vals))
        pv_cases.append(dict(kind='normalize', probs=list(probs), values=[list(v) if isinstance(v, tuple) else v for v in vals],
                             tuple_values=isinstance(vals[0], tuple), out_probs=pv.probs,
                             out_values=[list(v) if isinstance(v, tuple) else v for v in pv.values]))
    # nested ProbVal flattening (probVal.py:61-65)
    inner = ProbVal([.5, .5], [10, 20])
    outer = ProbVal([.25, .75], [inner, 30])
    pv_cases.append(dict(kind='nested', out_probs=outer.probs, out_values=outer.values))
    # fromUnzipped unwrapping
    pv_cases.append(dict(kind='unwrap', result=ProbVal.fromUnzipped([.5, .5], [7, 7])))
    # funcWrapper ordering: first ProbVal argument varies fastest (F10)
    a = ProbVal([.5, .5], [0, 1])
    b = ProbVal([.2, .3, .5], [10, 20, 30])
    r = funcWrapper(lambda x, c, y: (x, c, y), a, 'k', b)
    pv_cases.append(dict(kind='fanout', lens=[2, 3], out_probs=r.probs, out_values=[list(v) for v in r.values]))
    r2 = funcWrapper(lambda x, y: x + y, a, b)
    pv_cases.append(dict(kind='fanout_sum', out_probs=r2.probs, out_values=r2.values))
    # arithmetic dunders
    pv_cases.append(dict(kind='arith', expr='a*2+1', out_probs=(a * 2 + 1).probs, out_values=(a * 2 + 1).values))
    pv_cases.append(dict(kind='arith', expr='a+b', out_probs=(a + b).probs, out_values=(a + b).values))
    pv_cases.append(dict(kind='arith', expr='a==0', out_probs=(a == 0).probs, out_values=(a == 0).values))
    pv_cases.append(dict(kind='arith', expr='-b', out_probs=(-b).probs, out_values=(-b).values))
    with open(os.path.join(HERE, 'probval.json'), 'w') as f:
        json.dump(pv_cases, f, indent=0)

    # ---- DSL programs through the reference interpreter -------------------------------
    sout = record_scripts('scripts_src.json', 'scripts')
    # ---- random DSL programs (scripts/fuzz_dsl.py --emit), recorded the same way ----------
    fout = record_scripts('scripts_fuzz_src.json', 'scripts_fuzz')
    # ---- the same on 5-6 qubit registers (scripts/fuzz_dsl.py --big --emit: no swap, slot-aligned controls only, F5 / F6) ----
    bout = record_scripts('scripts_fuzz_big_src.json', 'scripts_fuzz_big')

    # ---- rc circuits through the reference's applyGate --------------------------------
    rarr = {}
    for (n, depth, seed) in ((4, 6, 4), (6, 6, 6), (7, 4, 7)):
        gates = rc(n, depth, seed)
        rho = np.zeros((1 << n, 1 << n), dtype=complex)
        rho[0, 0] = 1
        ref_built = 0
        for g in gates:
            if len(g.controls) == 0:
                u = rg.genGateForFullHilbertSpace(n, g.target, g.matrix())
                ref_built += 1
            elif slot_aligned(n, g.target, list(g.controls)):
                u = rg.genMultiControlledGate(n, list(g.controls), g.target, g.matrix())
                ref_built += 1
            else:
                u = orc.controlled_unitary(n, g.controls, g.target, g.matrix())
            rho = rg.applyGate(u, rho)
        rarr[f'rc_{n}_{depth}_{seed}'] = rho
        rarr[f'rc_{n}_{depth}_{seed}_info'] = np.array([len(gates), ref_built])
    np.savez_compressed(os.path.join(HERE, 'rc_small.npz'), **rarr)
    print('golden fixtures written:', len(meta), 'L1 cases,', len(sout), 'scripts,', len(fout), 'fuzz scripts,', len(bout), 'fuzz scripts on 5-6 qubits')


if __name__ == '__main__':
    main()
