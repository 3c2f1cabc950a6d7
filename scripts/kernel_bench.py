"""Achieved HBM bandwidth of every state-path kernel at the BASELINE config sizes (one GPU).
Times each entry point with CUDA events on the handle's stream (median of 5 after a warm-up) and
divides the kernel's algorithmic bytes by it.  Prints one JSON object."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from qbot_b200 import DeviceState, KET, DM
from qbot_b200.circuits import HADAMARD, PAULI_X, z_rot

PEAK = 6470.5
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    pass
out = {}


def timed(st, fn, reps=5):
    fn()
    st.sync()
    ts = []
    for _ in range(reps):
        st.timer_start()
        fn()
        ts.append(st.timer_stop())
    return float(np.median(ts))


_clock = DeviceState.zero_state(1)          # any handle: all handles share the device's compute stream


def wall(fn, reps=7):
    """(device ms, wall ms): CUDA events on the compute stream around the call (kernel time + gaps) and the
    host wall clock of the call including result-handle creation / destruction and any read-back"""
    fn()
    dev, ws = [], []
    for _ in range(reps):
        _clock.sync()
        _clock.timer_start()
        t0 = time.perf_counter()
        r = fn()
        w = 1e3 * (time.perf_counter() - t0)
        dev.append(_clock.timer_stop())
        ws.append(w)
        del r
    return float(np.median(dev)), float(np.median(ws))


def rec(name, ms, nbytes, note=''):
    wall_ms = None
    if isinstance(ms, tuple):
        ms, wall_ms = ms
    gbs = nbytes / ms / 1e6
    out[name] = {"ms": round(ms, 4), "bytes": int(nbytes), "GB/s": round(gbs, 1), "frac_of_peak": round(gbs / PEAK, 3), "note": note}
    if wall_ms is not None:
        out[name]["wall_ms"] = round(wall_ms, 4)
    print(f"{name:38s} {ms:9.3f} ms {gbs:8.1f} GB/s  {gbs / PEAK:5.2f}  " + (f"(call: {wall_ms:.3f} ms wall)  " if wall_ms is not None else '') + note, flush=True)


# ---- one-gate kernels on a 30-qubit ket (16 GiB) ---------------------------------------------------
n = 30
st = DeviceState.zero_state(n)
st.set_fusion(False)
S = 16 * (1 << n)
u2 = np.array([[0.6, 0.8j], [0.8j, 0.6]], dtype=complex)
rec('k_dense<1> H q29 (stride 1)', timed(st, lambda: (st.apply_gate(HADAMARD, n - 1), st.flush())), 2 * S)
rec('k_dense<1> U2 q0 (stride 2^29)', timed(st, lambda: (st.apply_gate(u2, 0), st.flush())), 2 * S)
rec('k_dense<1> CNOT c=3 t=17', timed(st, lambda: (st.apply_gate(PAULI_X, 17, [3]), st.flush())), S, 'touches half the ket')
rec('k_diag RZ q12', timed(st, lambda: (st.apply_gate(z_rot(0.3), 12), st.flush())), 2 * S)
u4 = np.linalg.qr(np.random.default_rng(0).normal(size=(4, 4)) + 1j * np.random.default_rng(1).normal(size=(4, 4)))[0]
rec('k_dense<2> U4 q10-11', timed(st, lambda: (st.apply_gate(u4, 10), st.flush())), 2 * S)
u8 = np.linalg.qr(np.random.default_rng(2).normal(size=(8, 8)) + 1j * np.random.default_rng(3).normal(size=(8, 8)))[0]
rec('k_dense<3> U8 q5-7', timed(st, lambda: (st.apply_gate(u8, 5), st.flush())), 2 * S)
rec('k_swap q2 <-> q20', timed(st, lambda: (st.apply_swap(2, 20), st.flush())), S, 'half the ket moves')
rec('k_bins probs of 4 qubits (ket)', wall(lambda: st.probs([0, 10, 20, 29])), S, 'incl. the 2^m read-back')
rec('k_bins norm (ket)', wall(lambda: st.norm2()), S)
rec('k_fill_basis', wall(lambda: DeviceState.zero_state(n)), S, 'allocation (recycled) + fill')
del st

# ---- config 4: 4096 branch kets of 16 qubits (4 GiB) ----------------------------------------------
B, nb = 4096, 16
rng = np.random.default_rng(4)
fac = rng.normal(size=(B, nb, 2)) + 1j * rng.normal(size=(B, nb, 2))
fac /= np.linalg.norm(fac, axis=-1, keepdims=True)
bs = DeviceState.product_batch(fac)
SB = 16 * B * (1 << nb)
rec('k_init_product 4096 x 16q', wall(lambda: DeviceState.product_batch(fac)), SB, 'incl. descriptor upload')
mats = np.stack([z_rot(t) for t in rng.uniform(0, 6, B)])
tg = [int(t) for t in rng.integers(0, nb, B)]
rec('k_dense_batched per-branch RZ', timed(bs, lambda: bs.apply_gate_batched(mats, tg)), 2 * SB)
xm = np.broadcast_to(PAULI_X, (B, 2, 2)).copy()
rec('k_dense_batched per-branch X', timed(bs, lambda: bs.apply_gate_batched(xm, tg)), 2 * SB)
rec('k_bins probs 4096 x 2^4', wall(lambda: bs.probs([1, 6, 11, 15])), SB, 'incl. the 512 KiB read-back')
del bs
# the weighted branch reduction is a density-matrix notion (sum_b p_b rho_b): 256 branches of a 10-qubit rho = 4 GiB
Bm, nm = 256, 10
bm = DeviceState.zero_state(nm, kind=DM).broadcast(Bm)
pr = rng.uniform(size=Bm)
pr /= pr.sum()
rec('k_mix_branches 256 x 10q DM -> 1', wall(lambda: bm.mix_branches(pr)), 16 * (Bm + 1) * (1 << (2 * nm)))
rec('k_broadcast 10q DM -> 256', wall(lambda: bm.broadcast(Bm) if False else DeviceState.zero_state(nm, kind=DM).broadcast(Bm)), 16 * Bm * (1 << (2 * nm)), 'incl. allocation')
del bm

# ---- config 3: 12-qubit density matrix (256 MiB) --------------------------------------------------
nd = 12
rho = DeviceState.zero_state(nd, kind=DM)
SD = 16 * (1 << (2 * nd))
rho.set_fusion(False)
rec('DM conjugation H (2 x k_dense<1>)', timed(rho, lambda: (rho.apply_gate(HADAMARD, 5), rho.flush())), 4 * SD)
rho.set_fusion(True)
rec('k_bins probs of 2 qubits (DM diagonal)', wall(lambda: rho.probs([2, 7])), 16 * (1 << nd), 'reads the diagonal only; latency-bound')
keep = [q for q in range(nd) if q not in (1, 4, 6, 10)]
rec('k_ptrace keep 8 of 12', wall(lambda: rho.ptrace_keep(keep)), 16 * (1 << (nd + 8)) + 16 * (1 << 16), 'reads the 2^(n+keep) entries whose traced row / column bits agree: 16 MiB of the 256 MiB')
rec('k_ptrace keep 2 of 12', wall(lambda: rho.ptrace_keep([3, 9])), 16 * (1 << (nd + 2)), '256 KiB read: latency-bound')
a = DeviceState.zero_state(2, kind=DM)
b8 = DeviceState.zero_state(10, kind=DM)
rec('k_scatter 2q (x) 10q -> 12q', wall(lambda: DeviceState.scatter_product(a, b8, [3, 9], [q for q in range(nd) if q not in (3, 9)])), SD)
r2 = rho.clone()
rec('k_mix 2 x 12q DM', wall(lambda: DeviceState.mix([0.5, 0.5], [rho, r2])), 3 * SD)
k12 = DeviceState.zero_state(nd)
from qbot_b200.host.interp import Interpreter
from qbot_b200.host.namespace import globalNameSpace as _gns
_measure = Interpreter(DeviceState).state_ops['measure']


def meas2():
    return _measure(rho, _gns['comp'], [3, 9], True)


rec('k_outer 12q ket -> DM', wall(lambda: k12.outer(True)), SD)
rec('measure 2 of 12 qubits (ptrace x2 + probs + collapse + scatter)', wall(lambda: meas2()), 3 * SD, 'whole meas op: reads rho twice (rho_A, rho_B), writes the new rho')
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out', 'kernel_bench.json'), 'w'), indent=1)
