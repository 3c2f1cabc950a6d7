"""Where the end-to-end time of the config-3 program (12-qubit density matrix, executeTxt) goes: cProfile of
the host side + wall clock per op kind."""
import os, sys, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import qbot_b200
from qbot_b200 import circuits

program = circuits.c3_program(12, 50, 12)
for _ in range(3):
    ns = qbot_b200.executeTxt(program)
    final = np.asarray(ns['state'])
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    ns = qbot_b200.executeTxt(program)
    final = np.asarray(ns['state'])
torch.cuda.synchronize()
print(f"executeTxt: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms per program ({len(program.splitlines())} lines)")
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    ns = qbot_b200.executeTxt(program)
    final = np.asarray(ns['state'])
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(25)
print(s.getvalue()[:6000])
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(30)
print(s.getvalue()[:7000])
