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


# ---------------------------------------------------------------------------------------------
# BASELINE config 3 and 4 workloads
# ---------------------------------------------------------------------------------------------
@dataclass
class C3Op:
    """One state operation of the config-3 program: kind 'gate' (a circuit gate), 'meas' (two
    qubits, computational basis, product-state collapse), 'pgate' (Hadamard on a ProbVal target)
    or 'disc' (final partial trace)."""
    kind: str
    gate: Gate = None
    qubits: Tuple[int, ...] = ()
    name: str = ''

    def dsl(self) -> str:
        if self.kind == 'gate':
            return self.gate.dsl()
        if self.kind == 'meas':
            return f"meas {self.name} ; comp ; {list(self.qubits)}"
        if self.kind == 'pgate':
            return f"gate hadamardGate ; ProbVal([.5,.5],[{self.qubits[0]},{self.qubits[1]}])"
        return f"disc {list(self.qubits)}"


def c3_ops(n: int = 12, depth: int = 50, seed: int = 12, every: int = 10) -> List[C3Op]:
    """Config 3: rho_0 = |0..0><0..0| on n qubits, rc(n, depth, seed) with, after every `every`
    layers, `meas m_i ; comp ; [two random qubits]` followed by one ProbVal-target gate
    `gate hadamardGate ; ProbVal([.5,.5],[a,b])`, and a final `disc` of min(4, n-1) random qubits.
    The extra choices come from their own generator (seed, 3), so the gate list is exactly rc()'s."""
    gates = rc(n, depth, seed)
    rng = np.random.default_rng([seed, 3])
    # rc() emits whole layers: cut the flat list back into layers by replaying its qubit budget
    ops: List[C3Op] = []
    used, layer, mi = 0, 0, 0
    for g in gates:
        ops.append(C3Op('gate', gate=g))
        used += 1 + len(g.controls)
        if used == n:
            used = 0
            layer += 1
            if layer % every == 0 and layer < depth:
                q = [int(x) for x in rng.choice(n, size=2, replace=False)]
                ops.append(C3Op('meas', qubits=tuple(q), name=f"m{mi}"))
                mi += 1
                ab = [int(x) for x in rng.choice(n, size=2, replace=False)]
                ops.append(C3Op('pgate', qubits=tuple(ab)))
    drop = sorted(int(x) for x in rng.choice(n, size=min(4, n - 1), replace=False))
    ops.append(C3Op('disc', qubits=tuple(drop)))
    return ops


def c3_program(n: int = 12, depth: int = 50, seed: int = 12, every: int = 10) -> str:
    """Config 3 as a qbot program (what a user of the reference would run through executeTxt)."""
    return "\n".join([f"qset tensorExp(comp[0], {n})"] + [op.dsl() for op in c3_ops(n, depth, seed, every)])


def c4_inputs(nbranch: int = 4096, n: int = 16, seed: int = 16):
    """Config 4: `nbranch` random product kets (per-qubit angles theta ~ U[0, pi), phi ~ U[0, 2 pi)),
    branch weights ~ U(0, 1] normalised, the shared circuit rc(n, 10, seed), a per-branch RZ(theta_b)
    on a per-branch target t_b, and 4 fixed measured qubits.
    Returns (factors [B, n, 2], weights [B], gates, rz_angles [B], rz_targets [B], measured qubits)."""
    rng = np.random.default_rng(seed)
    th = rng.uniform(0, np.pi, size=(nbranch, n))
    ph = rng.uniform(0, 2 * np.pi, size=(nbranch, n))
    factors = np.stack([np.cos(th / 2), np.exp(1j * ph) * np.sin(th / 2)], axis=-1)
    w = 1.0 - rng.uniform(0, 1, size=nbranch)          # (0, 1]
    w /= w.sum()
    gates = rc(n, 10, seed)
    ang = rng.uniform(0, 2 * np.pi, size=nbranch)
    tgt = rng.integers(0, n, size=nbranch)
    measured = sorted({1 % n, (3 * n) // 8, (11 * n) // 16, n - 1})
    return factors, w, gates, ang, tgt, measured
