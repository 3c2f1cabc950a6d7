"""GPU: the randomly generated DSL programs of tests/golden/scripts_fuzz.json (scripts/fuzz_dsl.py,
outputs recorded from the real reference) through the interpreter mirror on the CUDA backend.
Registers and reduced densities to 1e-12 relative, outcome order / symbols / error text exact,
outcome weights to 1e-12 (the reference rounds them to 15 decimals after its own summation order)."""
import pytest

from test_host_logic import check_script

pytestmark = pytest.mark.gpu


def test_fuzzed_scripts_on_device(golden):
    from qbot_b200 import DeviceState
    for rec in golden.scripts_fuzz:
        check_script(rec, golden.scripts_fuzz_arr, DeviceState, prob_tol=1e-12)

