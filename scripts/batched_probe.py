import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qbot_b200 import DeviceState
from qbot_b200.circuits import z_rot
B, nb = 4096, 16
rng = np.random.default_rng(4)
fac = rng.normal(size=(B, nb, 2)) + 1j * rng.normal(size=(B, nb, 2))
fac /= np.linalg.norm(fac, axis=-1, keepdims=True)
bs = DeviceState.product_batch(fac)
mats = np.stack([z_rot(t) for t in rng.uniform(0, 6, B)])
tg = [int(t) for t in rng.integers(0, nb, B)]
for _ in range(3):
    bs.apply_gate_batched(mats, tg)
bs.sync()
print(bs.probs([0])[0])
print(bs.norm2()[:2])
