"""GPU tests written AFTER the last GPU session of round 2 (the round's GPU minutes were spent): they have run on the numpy double
and against the oracle / the reference fixtures on the CPU, but not yet on a B200.  They live in the file that sorts last so that
`pytest -x` reaches them only after every test that has already passed on hardware.

  - config 3 at its full size (12 qubits, 447 ops) against the fixture recorded from the stock reference (tests/golden/c3_12.*)
  - `disc` on a ket-mode register above 13 qubits (Tr_rest psi psi^dagger straight from the amplitudes)
  - the 70 random DSL programs on 5-6 qubit registers recorded from the reference (tests/golden/scripts_fuzz_big.*)
  - the sharded register behind the DSL ops with the start map a fresh product register chooses for itself (the default;
    tests/test_gpu_sharded.py keeps running the same program from the identity map, as it last did on hardware)"""
import json
import os

import numpy as np
import pytest

from qbot_b200 import circuits
from conftest import close, GOLDEN
from test_configs_golden import C3_FULL_RTOL, _close_on_deviation
from test_host_logic import check_script

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def c3_full():
    return json.load(open(os.path.join(GOLDEN, 'c3_12.json')))['c3_12'], np.load(os.path.join(GOLDEN, 'c3_12.npz'))


def test_device_config3_full_size(c3_full):
    """config 3 as benchmarked (12 qubits, 256 MiB density matrix, 447 ops) against the REAL reference's output:
    every measurement's weights, the final 8-qubit register, and the 12-qubit register before the `disc`
    (diagonal, sampled rows, trace, purity)"""
    import qbot_b200
    meta, arr = c3_full
    prog = circuits.c3_program(12, 50, 12)
    ns = qbot_b200.executeTxt(prog)
    assert close(np.asarray(ns['state']), arr['c3_12_state'], C3_FULL_RTOL)
    assert _close_on_deviation(np.asarray(ns['state']), arr['c3_12_state'])
    for name, p in meta['probs'].items():
        assert np.allclose(ns[name].probs, p, rtol=0, atol=1e-12), name
    lines = prog.split("\n")
    assert lines[-1].startswith('disc')
    ns = qbot_b200.executeTxt("\n".join(lines[:-1]))
    rho = np.asarray(ns['state'])
    assert rho.shape == (4096, 4096)
    assert close(np.diag(rho), arr['c3_12_before_disc_diag'], C3_FULL_RTOL)
    assert close(rho[meta['rows']], arr['c3_12_before_disc_rows'], C3_FULL_RTOL)
    assert abs(np.trace(rho) - complex(*meta['before_disc_trace'])) < 1e-11
    assert abs(np.vdot(rho.conj().T, rho).real - meta['before_disc_purity']) < 1e-11


def test_disc_on_a_large_ket_register():
    """`disc` on a ket-mode register above 13 qubits: Tr_rest psi psi^dagger of the kept qubits straight from the
    amplitudes (qb_ptrace on a ket), then an ordinary density-matrix register -- against the oracle's ket path."""
    import io
    from contextlib import redirect_stdout
    import qbot_b200
    from test_host_logic import _disc_ket_case
    n, drop = 20, [0, 1, 2, 4, 5, 7, 8, 10, 11, 13, 14, 16, 17, 19]
    prog, rho, peek = _disc_ket_case(n, drop)
    ns = qbot_b200.executeTxt(prog)
    st = ns['state']
    assert st.kind == 1 and st.nq == n - len(drop)
    assert np.max(np.abs(np.asarray(st) - rho)) < 1e-12
    assert np.max(np.abs(np.array(ns['p'].probs) - peek)) < 1e-12
    buf = io.StringIO()
    with pytest.raises(SystemExit), redirect_stdout(buf):
        qbot_b200.executeTxt(f"qset tensorExp(comp.kets[0], {n})\ndisc [0, 1]\n")
    assert "at most 13 qubits" in buf.getvalue()


def test_fuzzed_scripts_on_5_and_6_qubits_on_device(golden):
    from qbot_b200 import DeviceState
    for rec in golden.scripts_fuzz_big:
        check_script(rec, golden.scripts_fuzz_big_arr, DeviceState, prob_tol=1e-12)


def test_sharded_register_through_the_dsl_with_the_chosen_start_map(tmp_path):
    from test_gpu_sharded import run_register_worker
    run_register_worker(tmp_path, '1')
