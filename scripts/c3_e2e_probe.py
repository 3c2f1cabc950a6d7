import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qbot_b200
from qbot_b200 import circuits
program = circuits.c3_program(12, 50, 12)
junk = torch.zeros(32 << 20, dtype=torch.float64, device='cuda')
for _ in range(4):
    ns = qbot_b200.executeTxt(program); final = np.asarray(ns['state'])
torch.cuda.synchronize()
def loop(flush, sync_before, keep):
    ts = []
    ns = None
    for _ in range(5):
        if flush: junk.add_(1)
        if sync_before: torch.cuda.synchronize()
        t0 = time.perf_counter()
        if keep:
            ns = qbot_b200.executeTxt(program)
        else:
            ns = None
            ns = qbot_b200.executeTxt(program)
        final = np.asarray(ns['state'])
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return [round(x, 2) for x in ts]
print('plain          ', loop(False, False, True))
print('sync before    ', loop(False, True, True))
print('flush + sync   ', loop(True, True, True))
print('drop old first ', loop(True, True, False))
