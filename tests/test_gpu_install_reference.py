"""GPU: the real drop-in.  `qbot_b200.install()` re-registers the six state ops inside the UNMODIFIED
reference (the pod's install under baseline/_ref, made by __graft_entry__.build()) with the CUDA
DeviceState as the register, then the reference's OWN 78 unit tests and the three README programs run
through the reference's own interpreter / evaluator / ProbVal -- not through this repo's mirror.
Skipped when baseline/_ref is absent."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'baseline', '_ref')
have_ref = os.path.isdir(os.path.join(REF, 'qbot', 'tests'))


def _run(code):
    p = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    return p.stdout


@pytest.mark.skipif(not have_ref, reason="baseline/_ref (reference install) not present")
def test_reference_unit_tests_on_the_cuda_backend():
    out = _run(r'''
import sys, unittest
sys.dont_write_bytecode = True
sys.path[:0] = [%r, %r]
import qbot_b200
from qbot_b200 import DeviceState
qbot_b200.install()
import qbot.operators as ops
assert ops.operations['gate'][0].__module__ == 'qbot_b200.host.ops'
import qbot
ns = qbot.executeTxt("qset tensorExp(comp[0], 2)\ngate hadamardGate ; 0\n")
assert isinstance(ns['state'], DeviceState), type(ns['state'])           # the register really lives on the device
import qbot.tests.unitTests as ut
res = unittest.TextTestRunner(verbosity=0).run(unittest.defaultTestLoader.loadTestsFromModule(ut))
print("RAN", res.testsRun, "FAIL", len(res.failures), "ERR", len(res.errors))
for _, tb in res.failures + res.errors:
    print(tb[-1500:])
st = DeviceState.zero_state(2)
print("LAUNCHES", st.stats())
sys.exit(0 if res.wasSuccessful() and res.testsRun >= 78 else 1)
''' % (REF, ROOT))
    assert 'RAN 78 FAIL 0 ERR 0' in out, out[-3000:]


README_PROGRAMS = {
    'superdense': ("cdef results ; []\ncdef index ; 0\n\nmark loop\nqset bell[0]\ngate pauliXGate ; 0 ; [] ; (index & 0b01) != 0\n"
                   "gate pauliZGate ; 0 ; [] ; (index & 0b10) != 0\nmeas result ; bell\npydo results.append(result.probs)\n"
                   "cdef index ; index + 1\ncjmp loop ; index < 4\n",
                   "[[1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0]]"),
    'kickback': ("cdef results ; []\n\nnote eiganValue is 1\nqset tensorProd(comp[0], hada[0])\njump checkPhase\n\nnote eiganValue is -1\n"
                 "qset tensorProd(comp[0], hada[1])\njump checkPhase\n\nhalt\n\nmark checkPhase\ngate hadamardGate ; 0\n"
                 "gate pauliXGate   ; 1 ; 0\ngate hadamardGate ; 0\nmeas tmp ; comp ; 0\n"
                 "pydo results.append(1 if np_isclose(tmp.probs[0], 1.0) else -1)\nretr\n", "[1, -1]"),
    'deutsch': ("cdef results ; []\n\nnote constant f\ncdef f ; lambda x: 1\njump check\n\nnote balanced f\ncdef f ; lambda x: x\njump check\n\n"
                "halt\n\nmark check\nqset tensorProd(comp[0], hada[1])\ngate hadamardGate ; 0\ngate simonsGate(2, f)\ngate hadamardGate ; 0\n"
                "meas tmp ; comp ; 0\npydo results.append(\"constant\" if np_isclose(tmp.probs[0], 1.0) else \"balanced\")\nretr\n",
                "['constant', 'balanced']"),
}


@pytest.mark.skipif(not have_ref, reason="baseline/_ref (reference install) not present")
@pytest.mark.parametrize('name', sorted(README_PROGRAMS))
def test_readme_programs_through_the_installed_reference(name):
    """BASELINE config 1 on the CUDA backend, and on the stock numpy path of the same checkout beside it"""
    text, want = README_PROGRAMS[name]
    out = _run(r'''
import sys
sys.dont_write_bytecode = True
sys.path[:0] = [%r, %r]
import qbot, qbot_b200
stock = qbot.executeTxt(%r)['results']
qbot_b200.install()
ours = qbot.executeTxt(%r)['results']
qbot_b200.uninstall()
again = qbot.executeTxt(%r)['results']
print("STOCK", [[float(x) for x in r] if isinstance(r, list) else r for r in stock])
print("OURS", [[float(x) for x in r] if isinstance(r, list) else r for r in ours])
assert stock == ours == again, (stock, ours, again)
''' % (REF, ROOT, text, text, text))
    assert ("OURS " + want) in out, out


@pytest.mark.skipif(not have_ref, reason="baseline/_ref (reference install) not present")
def test_large_ket_program_through_the_installed_reference():
    """rows f1 + b on the device: the UNMODIFIED reference's executeTxt drives a 22-qubit register (a 2^44-entry
    density matrix in its own representation): qset of a lazy product ket, `gate` lines fused into sweeps, `peek`;
    probabilities against the oracle's ket path"""
    out = _run(r'''
import sys
sys.dont_write_bytecode = True
sys.path[:0] = [%r, %r]
import numpy as np
import qbot, qbot_b200
from qbot_b200 import DeviceState, circuits
from qbot_b200.state import KET
from oracle import qbot_oracle as orc
qbot_b200.install()
n = 22
gates = circuits.rc(n, 4, 22)
qs = [0, 7, 14, 21]
script = "\n".join(["qset tensorExp(comp.kets[0], %%d)" %% n] + [g.dsl() for g in gates] + ["peek r ; comp ; %%s" %% qs])
ns = qbot.executeTxt(script)
st = ns['state']
assert isinstance(st, DeviceState) and st.nq == n and st.kind == KET, (type(st), st.nq)
psi = np.zeros(1 << n, dtype=complex)
psi[0] = 1
for g in gates:
    psi = orc.ket_apply(psi, n, g.target, g.matrix(), g.controls)
err = float(np.max(np.abs(np.array(ns['r'].probs) - orc.ket_probs(psi, n, qs))))
amp = float(np.max(np.abs(np.asarray(st) - psi)))
print("ERR", err, amp, st.stats()['fused_passes'])
assert err < 1e-12 and amp < 1e-12, (err, amp)
assert st.stats()['fused_passes'] >= 1
''' % (REF, ROOT))
    assert 'ERR' in out, out[-2000:]
