"""GPU: every DSL program of tests/golden/scripts.json (the reference's own test programs and
README examples, outputs recorded from the real reference) through the interpreter mirror on
the CUDA backend -- final register, measurement results, ProbVal-driven mixtures, control
flow and error text."""
import pytest

from test_host_logic import check_script

pytestmark = pytest.mark.gpu


def test_all_golden_scripts_on_device(golden):
    from qbot_b200 import DeviceState
    for rec in golden.scripts:
        check_script(rec, golden.scripts_arr, DeviceState)


def test_exact_equality_cases_stay_exact(golden):
    """The reference's tests assert np.array_equal on these programs' registers."""
    import numpy as np
    from qbot_b200 import DeviceState
    import qbot_b200
    by = {r['name']: r for r in golden.scripts}
    for name in ('gate_h', 'cgate_00', 'cgate_01_list', 'cgate_01_int', 'disc_plain', 'toffoli_slot', 'toffoli_upside_down'):
        ns = qbot_b200.executeTxt(by[name]['text'], state_cls=DeviceState)
        assert np.array_equal(np.asarray(ns['state']), golden.scripts_arr[by[name]['state']]), name
