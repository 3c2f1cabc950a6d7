"""Diagnostic: cost of one specialised sweep as a function of what it contains (30 qubits)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qbot_b200 import DeviceState
from qbot_b200.circuits import HADAMARD, PAULI_X, z_rot

n = 30
hi = [29, 27, 25, 23, 21, 19, 17]
lo = [4, 3, 2, 1, 0]
bits = hi + lo
q = lambda b: n - 1 - b
def H(b): return (HADAMARD, q(b), ())
def RZ(b, t=0.3): return (z_rot(t + 0.01 * b), q(b), ())
def CX(c, t): return (PAULI_X, q(t), (q(c),))
def CCX(c1, c2, t): return (PAULI_X, q(t), (q(c1), q(c2)))
cases = {
    '12 H': [H(b) for b in bits],
    '24 H (2 layers)': [H(b) for b in bits] * 2,
    '48 H (4 layers)': [H(b) for b in bits] * 4,
    '12 H + 12 RZ': [g for b in bits for g in (H(b), RZ(b))],
    '2 x (12 H + 12 RZ)': [g for b in bits for g in (H(b), RZ(b))] * 2,
    '4 x (12 H + 12 RZ)': [g for b in bits for g in (H(b), RZ(b))] * 4,
    '12 H + 11 CX chain': [H(b) for b in bits] + [CX(bits[i], bits[i + 1]) for i in range(11)],
    '12 H + 22 CX (chain both ways)': [H(b) for b in bits] + [CX(bits[i], bits[i + 1]) for i in range(11)] + [CX(bits[i + 1], bits[i]) for i in range(11)],
    '12 H + 10 CCX': [H(b) for b in bits] + [CCX(bits[i], bits[i + 1], bits[i + 2]) for i in range(10)],
    '12 H + 11 CX with outside controls': [H(b) for b in bits] + [CX(28 - 2 * (i % 6), bits[i]) for i in range(11)],
}
st = DeviceState.zero_state(n)
st.set_jit(2)
for name, gl in cases.items():
    for rep in range(2):
        for m, t, c in gl:
            st.apply_gate(m, t, c)
        st.flush()
    st.sync()
    st.reset_stats()
    st.timer_start()
    reps = 4
    for rep in range(reps):
        for m, t, c in gl:
            st.apply_gate(m, t, c)
        st.flush()
    ms = st.timer_stop()
    s = st.stats()
    print(f"{name:40s} gates {len(gl):3d} sweeps/rep {s['fused_passes'] / reps:.0f}  {ms / reps:7.3f} ms/rep  {ms / max(s['fused_passes'], 1):7.3f} ms/sweep")
