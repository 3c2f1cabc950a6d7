"""Diagnostic: how fast is one specialised sweep when it does almost no arithmetic, for contiguous
tiles and for tiles whose 512-byte runs are scattered by high tile bits?  (HBM-side floor of the
tile access pattern.)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qbot_b200 import DeviceState
from qbot_b200.circuits import HADAMARD, PAULI_X

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
cases = {
    'contiguous tile (X on index bit 5)': [(PAULI_X, n - 1 - 5)],
    'scattered: X on index bits 29,27,25,23,21,19,17': [(PAULI_X, n - 1 - b) for b in (29, 27, 25, 23, 21, 19, 17)],
    'scattered: X on index bits 12,11,10,9,8,7,6': [(PAULI_X, n - 1 - b) for b in (12, 11, 10, 9, 8, 7, 6)],
    'scattered + H on the same 7 high bits': [(HADAMARD, n - 1 - b) for b in (29, 27, 25, 23, 21, 19, 17)],
    'H on 7 high bits + H on 5 low bits': [(HADAMARD, n - 1 - b) for b in (29, 27, 25, 23, 21, 19, 17, 4, 3, 2, 1, 0)],
}
st = DeviceState.zero_state(n)
st.set_jit(2)
for name, gl in cases.items():
    for rep in range(3):
        for m, q in gl:
            st.apply_gate(m, q)
        st.flush()
    st.sync()
    st.reset_stats()
    st.timer_start()
    reps = 5
    for rep in range(reps):
        for m, q in gl:
            st.apply_gate(m, q)
        st.flush()
    ms = st.timer_stop()
    s = st.stats()
    per = ms / max(s['fused_passes'], 1)
    print(f"{name:55s} sweeps/rep {s['fused_passes'] / reps:.0f}  {per:7.3f} ms/sweep  {32 * 2**n / per / 1e6:7.1f} GB/s")
