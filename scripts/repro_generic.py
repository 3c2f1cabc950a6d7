import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qbot_b200 import DeviceState
from qbot_b200.circuits import rc
n = 14
gates = rc(n, 8, 3)
st = DeviceState.zero_state(n)
for rep in range(3):
    for g in gates:
        st.apply_gate(g.matrix(), g.target, g.controls)
    st.flush()
    print('rep', rep, 'ok', st.stats())
print(np.asarray(st)[:2])
