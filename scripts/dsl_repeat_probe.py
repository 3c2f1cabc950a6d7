"""Diagnostic: timing and executor of repeated executeTxt calls of the same program."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qbot_b200
from qbot_b200 import _lib
from qbot_b200.circuits import rc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
gates = rc(n, 20, n)
script = "\n".join([f"qset tensorExp(comp.kets[0], {n})"] + [g.dsl() for g in gates] + [f"peek r ; comp ; [0, 5, {n - 1}]"])
for rep in range(5):
    t0 = time.perf_counter()
    ns = qbot_b200.executeTxt(script)
    t1 = time.perf_counter()
    st = ns['state'].stats()
    print(rep, f"{1e3 * (t1 - t0):8.1f} ms", {k: st[k] for k in ('fused_passes', 'jit_passes', 'kernel_launches')}, _lib.jit_info())
    del ns

# finer: where does the time of one call go?
import gc
from qbot_b200.state import DeviceState
T = {}
def timed(cls, name):
    orig = getattr(cls, name)
    def w(*a, **k):
        t0 = time.perf_counter()
        try:
            return orig(*a, **k)
        finally:
            T[name] = T.get(name, 0.0) + time.perf_counter() - t0
    return w
for nm in ('probs', 'apply_gate', 'flush', '__del__'):
    setattr(DeviceState, nm, timed(DeviceState, nm))
orig_product = DeviceState.product.__func__
def product(cls, *a, **k):
    t0 = time.perf_counter()
    try:
        return orig_product(cls, *a, **k)
    finally:
        T['product'] = T.get('product', 0.0) + time.perf_counter() - t0
DeviceState.product = classmethod(product)
for rep in range(4):
    T.clear()
    t0 = time.perf_counter()
    ns = qbot_b200.executeTxt(script)
    t1 = time.perf_counter()
    del ns
    t2 = time.perf_counter()
    gc.collect()
    t3 = time.perf_counter()
    print('fine', rep, f"exec {1e3 * (t1 - t0):7.1f} del {1e3 * (t2 - t1):7.1f} gc {1e3 * (t3 - t2):7.1f} ms", {k: round(1e3 * v, 1) for k, v in T.items()})
