"""Differential fuzzer of the sharded register BEHIND THE DSL OPS (no GPU): random `.qb` programs on a 14-15 qubit product
ket -- gates (plain, controlled, dense 2-qubit, conditions), `swap`, `peek` in three bases, reads of rho_A, `disc` -- run

  (a) by two gloo ranks through qbot_b200.executeTxt with sharding enabled (ShardedRegister over numpy shards,
      qbot_b200/sharded_register.py: the chosen start map, exchanges, relabelling swaps, all-reduced weights), and
  (b) by one process on the ordinary ket-mode register of the numpy double, and
  (c) op by op by the ORACLE on the full ket (oracle/qbot_oracle.py: ket_apply, ket_swap, basis_weights, partial traces by
      reshaping), with the gate matrices from the reference's own namespace when /root/reference is there -- independent of
      host/ops.py, so the large-ket code paths that (a) and (b) share (measure_ket, the lazy rho_A, `disc` on a ket) are
      checked as well;

registers (gathered), outcome weights, reduced densities and the register left by `disc` must agree to 1e-12 on every rank.

    python scripts/fuzz_sharded_dsl.py --seeds 0:60

TEST INFRASTRUCTURE (numpy doubles; the oracle is behind them)."""
import argparse
import io
import os
import sys
from contextlib import redirect_stdout

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

ONE_Q = ['comp.kets[0]', 'comp.kets[1]', 'hadamard.kets[0]', 'hadamard.kets[1]']
G1 = ['hadamardGate', 'pauliXGate', 'pauliYGate', 'pauliZGate', 'xRotGate(%.3f)', 'yRotGate(%.3f)', 'zRotGate(%.3f)']


def program(seed):
    rng = np.random.default_rng(91_000 + seed)
    n = int(rng.integers(14, 16))
    parts = [ONE_Q[int(rng.integers(4))] for _ in range(n)]
    # (a run of equal factors as tensorExp, the rest listed: both constructors of the DSL)
    lines = ["qset tensorProd(%s)" % ", ".join(parts)] if rng.random() < 0.7 else [f"qset tensorExp({ONE_Q[int(rng.integers(4))]}, {n})"]
    names = []
    ops = [('init', lines[0][5:])]           # structured copy of the program for the oracle arm
    for _ in range(int(rng.integers(8, 40))):
        r = rng.random()
        if r < 0.62:
            g = G1[int(rng.integers(len(G1)))]
            g = g % rng.uniform(0, 6.28) if '%' in g else g
            k = 1
            if rng.random() < 0.12:
                g, k = ('qftGate(2)' if rng.random() < 0.5 else 'tensorProd(hadamardGate, pauliXGate)'), 2
            t = int(rng.integers(0, n - k + 1))
            free = [q for q in range(n) if not t <= q < t + k]
            cs = [int(x) for x in rng.choice(free, size=int(rng.choice([0, 0, 1, 1, 2])), replace=False)]
            ln = f"gate {g} ; {t} ; {cs}"
            on = True
            if rng.random() < 0.08:
                c = int(rng.integers(0, 4))
                ln += " ; %d > 1" % c
                on = c > 1
            lines.append(ln)
            if on:
                ops.append(('gate', g, t, cs))
        elif r < 0.72:
            a, b = rng.choice(n, 2, replace=False)
            lines.append(f"swap {a} ; {b}")
            ops.append(('swap', int(a), int(b)))
        else:
            basis = str(rng.choice(['comp', 'comp', 'hadamard', 'bell']))
            m = 2 if basis == 'bell' else int(rng.integers(1, 4))
            tg = [int(x) for x in rng.choice(n, m, replace=False)]
            name = f"m{len(names)}"
            names.append(name)
            lines.append(f"peek {name} ; {basis} ; {tg}")
            ops.append(('peek', name, basis, tg))
            if rng.random() < 0.3:
                lines.append(f"cdef rho_{name} ; np_array({name}.unMeasuredDensity)")
                names.append(f"rho_{name}")
                ops.append(('rho', f"rho_{name}", tg))
    disc = None
    if rng.random() < 0.4:
        keep = int(rng.integers(1, 9))
        disc = sorted(int(x) for x in rng.choice(n, n - keep, replace=False))
        lines.append(f"disc {disc}")
        ops.append(('disc', disc))
    return n, "\n".join(lines) + "\n", names, ops


def oracle_run(n, ops):
    """the program op by op on the full ket, by the oracle"""
    from oracle import qbot_oracle as orc
    if os.path.isdir('/root/reference'):
        sys.dont_write_bytecode = True
        if '/root/reference' not in sys.path:
            sys.path.insert(0, '/root/reference')
        import qbot.evaluation as ev              # the reference's own constants / constructors
        env = dict(ev.globalNameSpace) if hasattr(ev, 'globalNameSpace') else None
    else:
        env = None
    if env is None:
        from qbot_b200.host import namespace as nsm
        env = dict(nsm.build_namespace()) if hasattr(nsm, 'build_namespace') else {}
    def ev_(expr):
        return eval(expr, {'__builtins__': {}}, env)
    out = {}
    psi = None
    rho = None
    for op in ops:
        if op[0] == 'init':
            psi = np.asarray(ev_(op[1]), dtype=complex).reshape(-1)
        elif op[0] == 'gate':
            psi = orc.ket_apply(psi, n, op[2], np.asarray(ev_(op[1]), dtype=complex), op[3])
        elif op[0] == 'swap':
            psi = orc.ket_swap(psi, n, op[1], op[2])
        elif op[0] == 'peek':
            kets = ev_(op[2]).kets
            tg = sorted(set(op[3]))
            w = orc.basis_weights(psi, n, tg, kets)
            out[op[1]] = np.array([round(float(x), 15) for x in (w / w.sum())])
        elif op[0] == 'rho':
            keep = sorted(set(op[2]))
            rest = [q for q in range(n) if q not in keep]
            m = np.ascontiguousarray(psi.reshape([2] * n).transpose(keep + rest)).reshape(1 << len(keep), -1)
            out[op[1]] = m @ m.conj().T
        elif op[0] == 'disc':
            keep = [q for q in range(n) if q not in op[1]]
            m = np.ascontiguousarray(psi.reshape([2] * n).transpose(keep + list(op[1]))).reshape(1 << len(keep), -1)
            rho = m @ m.conj().T
    out['state'] = rho if rho is not None else psi
    return out


def collect(ns, names):
    out = {'state': np.asarray(ns['state'])}
    for v in names:
        x = ns[v]
        out[v] = np.asarray(x.probs, dtype=float) if hasattr(x, 'probs') else np.asarray(x)
    return out


def worker(rank, world, port, lo, hi, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch.distributed as dist
    import qbot_b200
    from qbot_b200 import sharded_register as sr
    from qbot_b200.sharded import TorchComm
    from np_shard import NumpyShard
    from fake_backend import FakeState
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sr.enable(TorchComm(), shard_factory=NumpyShard, min_qubits=14)
    res = {}
    for seed in range(lo, hi):
        n, text, names, _ = program(seed)
        try:
            with redirect_stdout(io.StringIO()):
                ns = qbot_b200.executeTxt(text, state_cls=FakeState)
            kinds = type(ns['state']).__name__
            res[seed] = ('ok', collect(ns, names), kinds)
            del ns
        except BaseException as e:      # noqa: BLE001
            res[seed] = ('error', f"{type(e).__name__}: {e}"[:300], None)
    q.put((rank, res))
    sr.disable()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--seeds', default='0:40')
    ap.add_argument('--world', type=int, default=2)
    a = ap.parse_args()
    lo, hi = (int(x) for x in a.seeds.split(':'))
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 200)
    procs = [ctx.Process(target=worker, args=(r, a.world, port, lo, hi, q)) for r in range(a.world)]
    for p in procs:
        p.start()
    # (b) the single ket-mode register, meanwhile
    import qbot_b200
    from fake_backend import FakeState
    want = {}
    for seed in range(lo, hi):
        n, text, names, _ = program(seed)
        try:
            with redirect_stdout(io.StringIO()):
                ns = qbot_b200.executeTxt(text, state_cls=FakeState)
            want[seed] = ('ok', collect(ns, names))
        except BaseException as e:      # noqa: BLE001
            want[seed] = ('error', f"{type(e).__name__}: {e}"[:300])
    # (c) the oracle, op by op
    bad_oracle = 0
    for seed in range(lo, hi):
        n, text, names, ops = program(seed)
        if want[seed][0] != 'ok':
            continue
        exp = oracle_run(n, ops)
        for k, y in want[seed][1].items():
            x = exp[k]
            if x.shape != y.shape or np.max(np.abs(x - y)) > 1e-12:
                bad_oracle += 1
                print(f"seed {seed}: single register vs ORACLE: {k} differs ({x.shape} vs {y.shape}, max {np.max(np.abs(x - y)) if x.shape == y.shape else -1})\n{text}", flush=True)
                break
    got = [q.get(timeout=3600) for _ in range(a.world)]
    for p in procs:
        p.join(timeout=120)
    bad = sharded = 0
    for seed in range(lo, hi):
        for rank, res in got:
            st, val, kind = res[seed]
            wst, wval = want[seed]
            why = None
            if st != wst:
                why = f"rank {rank}: {st} ({val if st == 'error' else ''}) vs single register {wst} ({wval if wst == 'error' else ''})"
            elif st == 'ok':
                for k in wval:
                    x, y = val[k], wval[k]
                    if x.shape != y.shape or np.max(np.abs(x - y)) > 1e-12:
                        why = f"rank {rank}: {k} differs ({x.shape} vs {y.shape}, max {np.max(np.abs(x - y)) if x.shape == y.shape else -1})"
                        break
            if why:
                bad += 1
                print(f"seed {seed}: {why}\n{program(seed)[1]}", flush=True)
                break
        sharded += got[0][1][seed][2] == 'ShardedRegister'
    print(f"seeds {lo}:{hi}: {bad} differences between the sharded and the single register, {bad_oracle} between the single register and the "
          f"oracle ({sharded} programs ended on a ShardedRegister, the others on the register `disc` left)")
    sys.exit(1 if bad or bad_oracle else 0)


if __name__ == '__main__':
    main()
