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
