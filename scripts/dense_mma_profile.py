"""One launch of k_dense_mma per block size (k = 6, 7, 8) on a 26-qubit ket, for an ncu capture:
ncu --set full --clock-control none --import-source on -k regex:k_dense_mma -o gpurun_out/dense_mma python scripts/dense_mma_profile.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qbot_b200 import DeviceState

n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
st = DeviceState.zero_state(n)
for k in (6, 7, 8):
    rng = np.random.default_rng(k)
    u = np.linalg.qr(rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k)))[0]
    st.apply_gate(u, 5)
    st.flush()
st.sync()
print("norm", float(st.probs([])[0]))
