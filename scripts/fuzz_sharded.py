"""CPU fuzzer of the sharded ket's host logic (no GPU): random circuits (rc + dense / diagonal / swap-like 2-qubit
blocks with controls) on P virtual ranks over numpy shards -- qubit map, exchange planning (plain and pipelined),
gate localisation, probability gather -- compared with the oracle's strided update on the full register.

    python scripts/fuzz_sharded.py --seeds 0:300

TEST INFRASTRUCTURE: uses oracle/ as the checker (through tests/test_sharded_host.py's harness)."""
import argparse
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


GENERATED = False          # --generated: the shards run their gate lists as planned sweeps of generated kernel text (g++)


def case(seed):
    from oracle import qbot_oracle as orc
    from qbot_b200.sharded import ShardedKet
    from np_shard import NumpyShard, VirtualComm
    from test_sharded_host import circuit_ops, expected_ket
    rng = np.random.default_rng(77_000 + seed)
    world = int(rng.choice([2, 4, 8]))
    g = world.bit_length() - 1
    pipelined = rng.random() < 0.4
    split = int(rng.integers(1, 4)) if pipelined else 0
    lo = max(g + 2, 4) if not pipelined else max(g + 2 + split, 12)       # pipelined plans need room for parked + tile bits
    n = int(rng.integers(lo, lo + 4))
    if GENERATED:
        pipelined, split = False, 0          # (sub-block sweeps are a CUDA-shard feature)
        n = g + int(rng.integers(12, 15))    # every flush must be plannable as fused sweeps: >= one 12-bit tile per shard
    depth = int(rng.integers(1, 9))
    ops = circuit_ops(n, depth, seed)
    cuts = sorted(int(c) for c in rng.integers(0, len(ops) + 1, size=int(rng.integers(0, 3))))
    probe = [int(q) for q in rng.choice(n, size=int(rng.integers(1, min(n, 4) + 1)), replace=False)]
    reps = int(rng.integers(1, 3))
    # half of the cases start from a random PRODUCT ket set through init_product (what `qset tensorExp(..)` does): the first
    # flush chooses which qubits start on the rank bits (QubitMap.choose_initial), or keeps the identity map (lazy off)
    product = rng.random() < 0.5
    lazy = bool(rng.random() < 0.75)
    swaps = [tuple(int(x) for x in rng.choice(n, size=2, replace=False)) for _ in range(int(rng.integers(0, 3)))] if product else []
    desc = dict(seed=seed, world=world, n=n, depth=depth, split=split, cuts=cuts, probe=probe, reps=reps, gates=len(ops),
                product=product, lazy=lazy, swaps=swaps)
    if product:
        factors = rng.normal(size=(n, 2)) + 1j * rng.normal(size=(n, 2))
        factors /= np.linalg.norm(factors, axis=1, keepdims=True)
        want = np.array([1.0 + 0j])
        for q in range(n):
            want = np.kron(want, factors[q])
        # `swap` before the gates only relabels: the caller's qubit q is the stored ket's qubit perm[q]; the check below
        # un-relabels the gathered ket, so the expected ket is simply the circuit on the relabelled qubits
        perm = list(range(n))
        for a, b in swaps:
            perm[a], perm[b] = perm[b], perm[a]
        inv = [perm.index(q) for q in range(n)]
        want = np.ascontiguousarray(want.reshape([2] * n).transpose(perm)).reshape(-1)     # the ket as the caller numbers it
        for m, t, cs in ops:
            want = orc.ket_apply(want, n, t, m, cs)
        del inv
    else:
        want = expected_ket(n, ops)
    for _ in range(reps - 1):
        for m, t, cs in ops:
            want = orc.ket_apply(want, n, t, m, cs)
    shared = VirtualComm.Shared(world)
    out = [None] * world
    errors = []

    def work(rank):
        try:
            kw = dict(split=split) if split else {}
            factory = NumpyShard
            if GENERATED:
                from test_sharded_host import _GeneratedCodeShard as factory
            sk = ShardedKet(n, VirtualComm(shared, rank), shard_factory=factory, **kw)
            if split:
                sk.min_first_phase = int(rng.integers(2, 8)) if rank < 0 else 6
            if product:
                sk.lazy_map = lazy
                sk.init_product(list(factors))
                for a, b in swaps:
                    sk.swap_qubits(a, b)
            for _ in range(reps):
                for i, (m, t, cs) in enumerate(ops):
                    if i in cuts:
                        sk.flush()
                    sk.apply_gate(m, t, cs)
                sk.flush()
            out[rank] = dict(ket=sk.gather(), probs=sk.probs(probe), norm=sk.norm2())
        except Exception as e:      # noqa: BLE001
            errors.append(e)
            shared.barrier.abort()

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    if errors:
        return 'error', desc, repr(errors[0])
    err = max(float(np.max(np.abs(o['ket'] - want))) for o in out)
    perr = max(float(np.max(np.abs(o['probs'] - orc.ket_probs(want, n, probe)))) for o in out)
    nerr = max(abs(float(o['norm']) - 1.0) for o in out)
    worst = max(err, perr, nerr)
    return ('ok' if worst < 1e-12 else 'MISMATCH'), desc, worst


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--seeds', default='0:100')
    ap.add_argument('--generated', action='store_true',
                    help="shards execute their local gate lists through the fusion planner and the specialiser's generated source "
                         "compiled with g++ (tests/jit_emu.py) instead of one numpy update per gate: host logic + planner + code "
                         "generator in one chain; seconds per case")
    a = ap.parse_args()
    global GENERATED
    GENERATED = a.generated
    lo, hi = (int(x) for x in a.seeds.split(':'))
    bad = 0
    for seed in range(lo, hi):
        status, desc, info = case(seed)
        if status != 'ok':
            bad += 1
            print(status, desc, info, flush=True)
    print(f"seeds {lo}:{hi}: {bad} failures")
    sys.exit(1 if bad else 0)


if __name__ == '__main__':
    main()
