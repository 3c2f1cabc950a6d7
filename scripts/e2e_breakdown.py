"""Where the end-to-end time of the 30-qubit benchmark program goes (one GPU): wall clock per phase of
executeTxt -- register creation, gate lines (host interpretation + queueing), flush (sweeps), peek."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import qbot_b200
from qbot_b200 import circuits, DeviceState
from qbot_b200.host.interp import Interpreter

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
gates = circuits.rc(n, 20, n)
qs = [0, n // 3, (2 * n) // 3, n - 1]
lines = [f"qset tensorExp(comp.kets[0], {n})"] + [g.dsl() for g in gates] + [f"peek r ; comp ; {qs}"]
script = "\n".join(lines)
for _ in range(3):
    ns = qbot_b200.executeTxt(script)
    del ns
torch.cuda.synchronize()


def phase_times():
    it = Interpreter(DeviceState)
    ns = {'state': np.array([], dtype=complex), '__updated_state': False, '__marks': dict(), '__prev_jump': -1}
    t = [time.perf_counter()]
    it.runtime(ns, lines[:1])
    t.append(time.perf_counter())
    torch.cuda.synchronize()
    t.append(time.perf_counter())
    for i in range(1, len(lines) - 1):
        tokens = [lines[i][:4].lower()] + [s.strip() for s in lines[i][4:].split(';') if s.strip()]
        it.operations['gate'][0](ns, lines, i, tokens)
    t.append(time.perf_counter())
    ns['state'].flush()
    t.append(time.perf_counter())
    torch.cuda.synchronize()
    t.append(time.perf_counter())
    tokens = [lines[-1][:4].lower()] + [s.strip() for s in lines[-1][4:].split(';') if s.strip()]
    it.operations['peek'][0](ns, lines, len(lines) - 1, tokens)
    t.append(time.perf_counter())
    del ns
    torch.cuda.synchronize()
    t.append(time.perf_counter())
    return np.diff(np.array(t)) * 1e3


names = ['qset (call)', 'qset (device done)', 'gate lines (host)', 'flush (call)', 'flush (device done)', 'peek', 'destroy']
acc = np.median(np.array([phase_times() for _ in range(5)]), axis=0)
for nme, v in zip(names, acc):
    print(f"{nme:24s} {v:8.3f} ms")
print(f"{'sum':24s} {acc.sum():8.3f} ms")
t0 = time.perf_counter()
for _ in range(3):
    ns = qbot_b200.executeTxt(script)
    p = ns['r'].probs
    del ns
torch.cuda.synchronize()
print(f"{'executeTxt':24s} {(time.perf_counter() - t0) / 3 * 1e3:8.3f} ms per call")
