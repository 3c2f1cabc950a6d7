"""Drop-in check: the six ops installed into the REAL reference's dispatch table, then the
reference's own 78 unit tests run unchanged.  Needs /root/reference (build container only) --
skipped elsewhere.  On CPU the numpy test double stands in for the device; the CUDA backend
itself is exercised by the gpu-marked tests through the in-repo interpreter mirror."""
import os
import subprocess
import sys

import pytest

REF = '/root/reference'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_reference_unit_tests_pass_with_backend_installed():
    code = r'''
import sys, unittest
sys.dont_write_bytecode = True
sys.path[:0] = [%r, %r, %r]
from fake_backend import FakeState
import qbot_b200.integration as integ
integ.install(state_cls=FakeState)
import qbot.operators as ops
assert ops.operations['gate'][0].__module__ == 'qbot_b200.host.ops'
import qbot.tests.unitTests as ut
res = unittest.TextTestRunner(verbosity=0).run(unittest.defaultTestLoader.loadTestsFromModule(ut))
print("RAN", res.testsRun, "FAIL", len(res.failures), "ERR", len(res.errors))
sys.exit(0 if res.wasSuccessful() and res.testsRun >= 78 else 1)
''' % (REF, ROOT, os.path.join(ROOT, 'tests'))
    p = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert 'RAN 78 FAIL 0 ERR 0' in p.stdout


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_installed_normalize_equals_the_stock_pairwise_loop():
    """row f4: install() swaps the O(B^2) de-duplication of the REAL qbot.probVal.ProbVal for a hashed one
    on exact-equality values; results identical to the stock class, 4096 gate arrays in well under a second"""
    code = r'''
import sys, time
sys.dont_write_bytecode = True
sys.path[:0] = [%r, %r, %r]
import numpy as np
from fake_backend import FakeState
import qbot.probVal as pv
rng = np.random.default_rng(4)
cases = {
  'ints': [int(x) for x in rng.integers(0, 9, 300)],
  'arrays': [np.array([[1, 0], [0, np.exp(1j * float(k))]]) for k in rng.integers(0, 7, 200)],
  'strs': [str(int(x)) for x in rng.integers(0, 5, 100)],
  'floats': [float(x) for x in rng.integers(0, 5, 100)],
}
probs = {k: list(rng.uniform(0.01, 1, len(v))) for k, v in cases.items()}
for k in probs: probs[k][3] = 1e-7
stock = {k: pv.ProbVal(list(probs[k]), list(v)) for k, v in cases.items()}
import qbot_b200.integration as integ
integ.install(state_cls=FakeState)
for k, v in cases.items():
    got = pv.ProbVal(list(probs[k]), list(v))
    assert got.probs == stock[k].probs, k
    assert len(got.values) == len(stock[k].values) and all(pv.valsClose(a, b) for a, b in zip(got.values, stock[k].values)), k
big = [np.diag([1, np.exp(1j * (i %% 64))]) for i in range(4096)]
t0 = time.perf_counter()
r = pv.ProbVal([1 / 4096] * 4096, big)
dt = time.perf_counter() - t0
assert len(r.values) == 64 and dt < 2.0, (len(r.values), dt)
integ.uninstall()
assert pv.ProbVal.normalize is not None and pv.ProbVal([.5, .5], [1, 1]).probs == [1.0]
print("OK", dt)
''' % (REF, ROOT, os.path.join(ROOT, 'tests'))
    p = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_large_ket_program_through_the_real_reference():
    """rows f1 + b: with the ops installed, the UNMODIFIED reference's executeTxt runs a register it cannot
    represent itself -- `qset tensorExp(comp.kets[0], 16)` stays a product descriptor (install() swaps the
    namespace's tensorExp / tensorProd for the lazy ones), the register is a ket, `gate` lines and `peek` drive it;
    probabilities against the oracle's ket path"""
    code = r'''
import sys
sys.dont_write_bytecode = True
sys.path[:0] = [%r, %r, %r]
import numpy as np
from fake_backend import FakeState, KET
import qbot_b200.integration as integ
integ.install(state_cls=FakeState)
import qbot
from qbot_b200 import circuits
from oracle import qbot_oracle as orc
n = 16
gates = circuits.rc(n, 3, 5)
qs = [0, 5, 9, 15]
script = "\n".join(["qset tensorExp(comp.kets[0], %%d)" %% n] + [g.dsl() for g in gates] + ["peek r ; comp ; %%s" %% qs])
ns = qbot.executeTxt(script)
assert isinstance(ns['state'], FakeState) and ns['state'].kind == KET and ns['state'].nq == n, type(ns['state'])
psi = np.zeros(1 << n, dtype=complex)
psi[0] = 1
for g in gates:
    psi = orc.ket_apply(psi, n, g.target, g.matrix(), g.controls)
err = float(np.max(np.abs(np.array(ns['r'].probs) - orc.ket_probs(psi, n, qs))))
assert err < 1e-12, err
integ.uninstall()
small = qbot.executeTxt("cdef x ; tensorExp(comp.kets[0], 2)\n")['x']
assert isinstance(small, np.ndarray) and small.shape == (4,)      # the stock constructor is back
print("OK", err)
''' % (REF, ROOT, os.path.join(ROOT, 'tests'))
    p = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
