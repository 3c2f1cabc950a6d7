"""Static SASS statistics of the specialised sweeps of a circuit, without a GPU: plan the circuit as
the engine does, NVRTC-compile every sweep for sm_100a (qb_jit_check) and count the instructions of
each cubin by opcode class (cuobjdump -sass).  The per-sweep time model fitted to the ncu launch list
 is  t = max(5.4, 2.6 + 0.45 * stages + 0.0012 * instructions) ms  at 30 qubits,
so the instruction count is the figure of merit for changes to the code generator.

    python scripts/sass_count.py [--n 30 --depth 20 --seed 30] [--keep DIR]
"""
import argparse
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

CLASSES = [('fp64', r'^(DADD|DMUL|DFMA)'), ('lds/sts', r'^(LDS|STS)'), ('ldg/stg', r'^(LDG|STG|LD\.|ST\.|LDC|ULDC)'),
           ('bulk/mbar', r'^(UBLKCP|SYNCS|UTMA|ARRIVES|FENCE|MEMBAR)'), ('bar', r'^(BAR|WARPSYNC)'),
           ('mov/sel', r'^(MOV|SEL|FSEL|PRMT|SHFL|UMOV|R2UR|S2R|S2UR|CS2R)'),
           ('int', r'^(IMAD|IADD|LEA|SHF|LOP|ISETP|UIADD|ULEA|USHF|ULOP|UISETP|UIMAD|PLOP|POPC|FLO|I2F|F2I|VIADD|IABS|UFLO|USEL|UPLOP|UPOPC|BMSK|SGXT)'),
           ('branch', r'^(BRA|EXIT|BSSY|BSYNC|CALL|RET|NOP|WARPSYNC|YIELD|BREAK|JMP|BRX|NANOSLEEP)')]


def count(cubin):
    txt = subprocess.run(['cuobjdump', '-sass', cubin], capture_output=True, text=True, check=True).stdout
    c = collections.Counter()
    for line in txt.splitlines():
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if not m:
            continue
        op = m.group(1)
        for name, pat in CLASSES:
            if re.match(pat, op):
                c[name] += 1
                break
        else:
            c['other:' + op.split('.')[0]] += 1
        c['total'] += 1
    return c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=30)
    ap.add_argument('--depth', type=int, default=20)
    ap.add_argument('--seed', type=int, default=30)
    ap.add_argument('--keep')
    a = ap.parse_args()
    import plan_emu
    from qbot_b200 import _lib
    from qbot_b200.circuits import rc
    gl = plan_emu.circuit_to_bits(a.n, rc(a.n, a.depth, a.seed))
    d = a.keep or tempfile.mkdtemp(prefix='qb_sass_')
    os.makedirs(d, exist_ok=True)
    k = _lib.jit_check(a.n, gl, d)
    tot = collections.Counter()
    names = [n for n, _ in CLASSES]
    print('%-18s %7s ' % ('kernel', 'total') + ' '.join('%9s' % n for n in names) + '  other')
    for f in sorted(os.listdir(d)):
        if not f.endswith('.cubin'):
            continue
        c = count(os.path.join(d, f))
        tot.update(c)
        other = {k2: v for k2, v in c.items() if k2.startswith('other:')}
        print('%-18s %7d ' % (f, c['total']) + ' '.join('%9d' % c[n] for n in names) + '  ' + ' '.join('%s=%d' % (k2[6:], v) for k2, v in sorted(other.items())))
    other = {k2: v for k2, v in tot.items() if k2.startswith('other:')}
    print('%-18s %7d ' % ('sum of %d' % k, tot['total']) + ' '.join('%9d' % tot[n] for n in names) + '  ' + ' '.join('%s=%d' % (k2[6:], v) for k2, v in sorted(other.items())))


if __name__ == '__main__':
    main()
