"""Row f3: what a dense k-qubit block (k = 5..8: qftGate(k), user unitaries) costs through the existing
kernels (k_dense<5> in registers, k_big out of place) on a 26-qubit ket, against the two bounds that
apply: HBM (32 * 2^n bytes) and FP64 ((8 * 2^k - 2) * 2^n flops at the vector FP64 peak)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qbot_b200 import DeviceState
from qbot_b200.host import hostmath as hm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
st = DeviceState.zero_state(n)
st.apply_gate(hm.qft(4), 3)
st.sync()
out = {}
for k in (4, 5, 6, 7, 8):
    rng = np.random.default_rng(k)
    u = np.linalg.qr(rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k)))[0]
    for name, m in (('random unitary', u), ('qftGate', hm.qft(k))):
        for t in (n - k, 5):           # lowest (stride-1) block and a middle block
            st.apply_gate(m, t)
            st.sync()
            st.reset_stats()
            st.timer_start()
            for _ in range(3):
                st.apply_gate(m, t)
                st.flush()
            ms = st.timer_stop() / 3
            s = st.stats()
            flops = (8 * (1 << k) - 2) * (1 << n)
            rec = {"ms": round(ms, 3), "GB/s": round(32 * (1 << n) / ms / 1e6, 1), "TFLOP/s": round(flops / ms / 1e9, 2),
                   "launches_per_gate": s['kernel_launches'] / 3, "passes_per_gate": s['state_passes'] / 3}
            out[f"k={k} {name} first_target={t}"] = rec
            print(f"k={k} {name:15s} t={t:2d}  {rec}", flush=True)
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out', 'dense_big_probe.json'), 'w'), indent=1)
