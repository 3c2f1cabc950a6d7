"""Seeded synthetic circuits for the benchmark and parity configurations.

``rc(n, depth, seed)`` is the generator SURVEY.md section 8(d) specifies: start from |0...0>,
and for each of ``depth`` layers shuffle the n qubits and walk the shuffled list drawing a
gate type with p = {H .35, RZ(theta) .35, CNOT .20, Toffoli .10}; a gate consumes 1/1/2/3
qubits of the walk (the last one is the target, the others are controls) and falls back to H
when too few qubits remain.  Gates are emitted as the arguments of the reference's ``gate``
op (``qbot/operators.py:274-329``): (matrix name, firstTarget, controls[, theta]).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np

_S = 2 ** (-1 / 2)
# same values as the reference's namespace constants (qbot/evaluation.py:19-35)
HADAMARD = _S * np.array([[1, 1], [1, -1]], dtype=complex)
PAULI_X = np.array([[0, 1], [1, 0]], dtype=complex)


def z_rot(theta: float) -> np.ndarray:
    """diag(e^{-i theta/2}, e^{+i theta/2}) -- the reference's zRotGate (qbot/qgates.py:56-60)."""
    return np.array([[np.exp(-1j * theta / 2), 0], [0, np.exp(1j * theta / 2)]], dtype=complex)


@dataclass
class Gate:
    name: str                 # 'H' | 'RZ' | 'X'
    target: int               # reference qubit number (0 = most significant)
    controls: Tuple[int, ...] = ()
    theta: float = 0.0

    def matrix(self) -> np.ndarray:
        if self.name == 'H':
            return HADAMARD
        if self.name == 'X':
            return PAULI_X
        return z_rot(self.theta)

    def dsl(self) -> str:
        """The DSL line a user of the reference would write for this gate."""
        if self.name == 'H':
            return f"gate hadamardGate ; {self.target}"
        if self.name == 'RZ':
            return f"gate zRotGate({self.theta!r}) ; {self.target}"
        return f"gate pauliXGate ; {self.target} ; {list(self.controls)}"

    def algorithmic_bytes(self, n: int) -> int:
        """read + write of the touched amplitudes: 32 * 2^(n - #controls) B."""
        return 32 * (1 << (n - len(self.controls)))


def rc(n: int, depth: int, seed: int) -> List[Gate]:
    rng = np.random.default_rng(seed)
    gates: List[Gate] = []
    for _ in range(depth):
        order = [int(q) for q in rng.permutation(n)]
        pos = 0
        while pos < n:
            u = rng.random()
            left = n - pos
            if u < 0.35:
                kind, need = 'H', 1
            elif u < 0.70:
                kind, need = 'RZ', 1
            elif u < 0.90:
                kind, need = 'CNOT', 2
            else:
                kind, need = 'TOFF', 3
            theta = float(rng.uniform(0.0, 2 * np.pi))    # always drawn: keeps the stream aligned
            if need > left:
                kind, need = 'H', 1
            qs = order[pos:pos + need]
            pos += need
            if kind == 'H':
                gates.append(Gate('H', qs[0]))
            elif kind == 'RZ':
                gates.append(Gate('RZ', qs[0], (), theta))
            else:
                gates.append(Gate('X', qs[-1], tuple(qs[:-1])))
    return gates


def rc_script(n: int, depth: int, seed: int) -> str:
    """The circuit as a qbot program (state must already hold n qubits)."""
    return "\n".join(g.dsl() for g in rc(n, depth, seed))


def total_algorithmic_bytes(gates: List[Gate], n: int) -> int:
    return sum(g.algorithmic_bytes(n) for g in gates)
