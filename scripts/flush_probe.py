"""Diagnostic: wall vs GPU time of flush and probs on fresh registers (mode-1 specialisation)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qbot_b200 import DeviceState, KET
from qbot_b200.circuits import rc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
gates = rc(n, 20, n)
mats = [g.matrix() for g in gates]
zero = [np.array([1, 0], dtype=complex)] * n
for rep in range(8):
    t0 = time.perf_counter()
    st = DeviceState.product(zero, KET)
    t1 = time.perf_counter()
    for g, m in zip(gates, mats):
        st.apply_gate(m, g.target, g.controls)
    t2 = time.perf_counter()
    st.timer_start()
    st.flush()
    gpu_ms = st.timer_stop()
    t3 = time.perf_counter()
    p = st.probs([0, 5, n - 1])
    t4 = time.perf_counter()
    s = st.stats()
    del st
    t5 = time.perf_counter()
    print(rep, f"create {1e3*(t1-t0):6.1f} queue {1e3*(t2-t1):6.1f} flush wall {1e3*(t3-t2):7.1f} gpu {gpu_ms:7.1f} probs {1e3*(t4-t3):6.1f} del {1e3*(t5-t4):6.1f}",
          s['jit_passes'], s['fused_passes'])
