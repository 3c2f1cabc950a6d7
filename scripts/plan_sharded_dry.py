"""Dry run of the sharded ket's planner (host logic only, no GPU, no amplitudes): what a repeated circuit costs
per step on P ranks -- exchanges, flush segments, and the fused sweeps the C++ planner cuts each segment into
(tests/plan_emu.py, the specialiser's 32-amplitudes-per-thread plan shape at the engine's search effort).

    python scripts/plan_sharded_dry.py [--world 8 --qubits 34 --depth 10 --steps 16] [--fresh] [--identity-start]

--fresh           every step starts from a fresh product register (what the DSL e2e leg of bench_multi.py runs);
                  default: the circuit is applied again and again to one register (the device-resident `value` leg)
--identity-start  fresh registers start from the identity qubit map instead of QubitMap.choose_initial

Used to find that rc(34, 10) from the identity map needs a second exchange for its last gate on 8 ranks (DESIGN.md
section 5), and to check that the qubit map of the repeated circuit settles into a short cycle well inside the
bench's warm-up (no NVRTC compile can land in the timed steps).  TEST / ANALYSIS INFRASTRUCTURE."""
import argparse
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


class OneRank:
    def __init__(self, rank, world):
        self.rank, self.world = rank, world

    def barrier(self):
        pass


class RecordingShard:
    """what a shard would be asked to do"""
    supports_split = False

    def __init__(self, nl, comm):
        self.nl, self.cur, self.segments, self.exchanges = nl, [], [], 0

    def init_basis(self, has_one, local_index=0):
        pass

    def init_product(self, local_factors, coeff):
        pass

    def apply(self, m, tpos, cmask):
        self.cur.append((np.array(m), list(tpos), int(cmask)))

    def flush(self):
        if self.cur:
            self.segments.append(self.cur)
        self.cur = []

    def do_exchange(self, ex):
        self.exchanges += 1
        self.segments.append('X')

    def sync(self):
        pass


_sweeps = {}


def sweeps_of(nl, seg):
    import jit_emu
    key = hashlib.sha1(repr(nl).encode() + b''.join(m.tobytes() + repr((t, c)).encode() for m, t, c in seg)).hexdigest()
    if key not in _sweeps:
        _sweeps[key] = len(jit_emu.plan(nl, seg))
    return _sweeps[key], key


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--world', type=int, default=8)
    ap.add_argument('--qubits', type=int, default=34)
    ap.add_argument('--depth', type=int, default=10)
    ap.add_argument('--steps', type=int, default=16)
    ap.add_argument('--rank', type=int, default=1)
    ap.add_argument('--fresh', action='store_true')
    ap.add_argument('--identity-start', action='store_true')
    a = ap.parse_args()
    os.environ.setdefault('QBOT_B200_PLAN_R', '5')
    os.environ.setdefault('QBOT_B200_PLAN_TRIALS', '128')
    from qbot_b200 import circuits
    from qbot_b200.sharded import ShardedKet
    n = a.qubits
    gates = circuits.rc(n, a.depth, n)
    zero = [np.array([1, 0], dtype=complex)] * n
    sk = ShardedKet(n, OneRank(a.rank, a.world), shard_factory=RecordingShard)
    sk.lazy_map = not a.identity_start
    seen = set()
    for step in range(a.steps):
        if a.fresh:
            sk.init_product(zero)
        sk.shard.segments = []
        for g in gates:
            sk.apply_gate(g.matrix(), g.target, g.controls)
        sk.flush()
        row, total, fresh_keys = [], 0, 0
        for seg in sk.shard.segments:
            if isinstance(seg, str):
                row.append('X')
                continue
            s, key = sweeps_of(sk.map.nl, seg)
            fresh_keys += key not in seen
            seen.add(key)
            total += s
            row.append(f"{len(seg)} gates / {s} sweeps")
        nex = sum(1 for s in sk.shard.segments if isinstance(s, str))
        print(f"step {step:2d}: {total} sweeps, {nex} exchange(s), {fresh_keys} local gate list(s) not seen before   [{' | '.join(row)}]")


if __name__ == '__main__':
    main()
