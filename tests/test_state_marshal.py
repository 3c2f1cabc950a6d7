"""CPU: host-side marshalling of a batched launch (qbot_b200.state.marshal_batched) against the
per-branch definition: target index bits, control masks (rectangular and ragged), enable flags."""
import numpy as np
import pytest

from qbot_b200.state import marshal_batched


def reference(nq, mats, qubits, controls):
    tb = [nq - 1 - int(q) for qs in qubits for q in qs]
    masks = []
    for cs in (controls if controls is not None else [[]] * len(qubits)):
        cm = 0
        for c in cs:
            cm |= 1 << (nq - 1 - int(c))
        masks.append(cm)
    return tb, masks


@pytest.mark.parametrize('k', [1, 2])
def test_matches_per_branch_definition(k):
    rng = np.random.default_rng(k)
    nq, b = 40, 257
    mats = rng.normal(size=(b, 1 << k, 1 << k)) + 1j * rng.normal(size=(b, 1 << k, 1 << k))
    qubits = [[int(q) for q in rng.choice(nq, k, replace=False)] for _ in range(b)]
    rect = [[int(c) for c in rng.choice(nq, 2, replace=False)] for _ in range(b)]
    ragged = [[int(c) for c in rng.choice(nq, int(rng.integers(0, 4)), replace=False)] for _ in range(b)]
    for controls in (None, rect, ragged, [[] for _ in range(b)]):
        m, kk, tb, masks, en = marshal_batched(nq, b, mats, qubits, controls, None)
        wt, wm = reference(nq, mats, qubits, controls)
        assert kk == k and m.dtype == np.complex128 and m.flags.c_contiguous and np.array_equal(m, mats)
        assert tb.dtype == np.int32 and tb.tolist() == wt
        assert masks.dtype == np.uint64 and [int(x) for x in masks[:b]] == wm
        assert en is None
    en = marshal_batched(nq, b, mats, qubits, None, [i % 3 == 0 for i in range(b)])[4]
    assert en.dtype == np.uint8 and en.tolist() == [int(i % 3 == 0) for i in range(b)]


def test_high_control_bits_and_errors():
    nq, b = 64, 2
    mats = np.stack([np.eye(2, dtype=complex)] * b)
    m, k, tb, masks, _ = marshal_batched(nq, b, mats, [[5], [63]], [[0], [1]], None)
    assert [int(x) for x in masks] == [1 << 63, 1 << 62] and tb.tolist() == [58, 0]
    with pytest.raises(ValueError):
        marshal_batched(nq, b, mats, [[5, 6], [7, 8]])
    with pytest.raises(ValueError):
        marshal_batched(nq, b, mats, [[5], [7, 8]])
    with pytest.raises(ValueError):
        marshal_batched(nq, 3, mats, [[5], [7], [1]])
    with pytest.raises(ValueError):
        marshal_batched(nq, b, mats, [[5]])
