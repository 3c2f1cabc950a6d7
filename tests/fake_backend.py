"""TEST DOUBLE: a numpy/oracle-backed stand-in for ``qbot_b200.DeviceState``.

Lets the host logic (ops, interpreter mirror, ProbVal fan-out, validation, error text) be
checked on a machine without a GPU.  It is injected explicitly by tests
(``qbot_b200.executeTxt(text, state_cls=FakeState)``); the product never imports it and has
no CPU fallback.
"""
import numpy as np

from oracle import qbot_oracle as orc

KET, DM = 0, 1


class FakeState:
    _qb_device_state = True
    __array_priority__ = 1000
    __hash__ = None

    def __init__(self, data, kind, nq, nbranch=1):
        self.data = np.array(data, dtype=complex)
        self.kind, self.nq, self.nbranch = kind, nq, nbranch

    @classmethod
    def from_host(cls, a, device=0):
        a = np.asarray(a, dtype=complex)
        if a.ndim == 2:
            return cls(a, DM, orc.ilog2(a.shape[0]))
        return cls(a, KET, orc.ilog2(a.shape[0]))

    @classmethod
    def product(cls, factors, kind=KET, device=0):
        out = np.asarray(factors[0], dtype=complex)
        for f in factors[1:]:
            out = np.kron(out, np.asarray(f, dtype=complex))
        return cls(out, kind, len(factors))

    def clone(self):
        return FakeState(self.data.copy(), self.kind, self.nq, self.nbranch)

    @property
    def shape(self):
        return self.data.shape

    @property
    def size(self):
        return self.data.size

    @property
    def ndim(self):
        return self.data.ndim

    def __array__(self, dtype=None, copy=None):
        return self.data

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError(name)
        return getattr(self.data, name)

    def __getitem__(self, i):
        return self.data[i]

    def __eq__(self, other):
        return self.data == np.asarray(other)

    def _apply_one(self, arr, matrix, qubits, controls):
        m = np.asarray(matrix, dtype=complex)
        k = len(qubits)
        n = self.nq
        # arbitrary qubit list: permute to a contiguous block through the definitional matrix
        if list(qubits) == list(range(qubits[0], qubits[0] + k)):
            if self.kind == DM:
                return orc.dm_apply(arr, n, qubits[0], m, controls)
            return orc.ket_apply(arr, n, qubits[0], m, controls)
        assert k == 2 and not controls, "fake backend: only 2-qubit non-contiguous gates (swap)"
        u = orc.swap_unitary(n, qubits[0], qubits[1])
        return orc.conjugate(u, arr) if self.kind == DM else u @ arr

    def apply_gate(self, matrix, first_target=0, controls=()):
        k = orc.ilog2(np.asarray(matrix).shape[0])
        self.data = self._apply_one(self.data, matrix, list(range(first_target, first_target + k)), list(controls))
        return self

    def apply_swap(self, a, b):
        if self.kind == KET:                  # (no 2^n x 2^n permutation matrix for the large ket-mode registers)
            self.data = orc.ket_swap(self.data, self.nq, a, b)
            return self
        u = orc.swap_unitary(self.nq, a, b)
        self.data = orc.conjugate(u, self.data) if self.kind == DM else u @ self.data
        return self

    def broadcast(self, nbranch):
        return FakeState(np.stack([self.data] * nbranch), self.kind, self.nq, nbranch)

    def apply_gate_batched_qubits(self, matrices, target_qubits, controls=None, enable=None):
        for b in range(self.nbranch):
            if enable is not None and not enable[b]:
                continue
            self.data[b] = self._apply_one(self.data[b], matrices[b], list(target_qubits[b]),
                                           list(controls[b]) if controls is not None else [])
        return self

    def apply_branch_gates(self, items):
        for b, it in enumerate(items):
            if it is not None:
                self.data[b] = self._apply_one(self.data[b], it[0], list(it[1]), list(it[2]))
        return self

    def mix_branches(self, probs):
        return FakeState(orc.ensemble(list(probs), list(self.data)), self.kind, self.nq)

    @staticmethod
    def mix(probs, states):
        if len(probs) != len(states):
            raise Exception("number of state vectors an number of probabilites must equal")
        for s in states:
            if s.shape != states[0].shape:
                raise ValueError("operands could not be broadcast together")
        return FakeState(orc.ensemble(list(probs), [s.data for s in states]), states[0].kind, states[0].nq)

    def ptrace_keep(self, keep):
        keep = list(keep)
        if self.kind == KET:          # Tr_rest |psi><psi| straight from the amplitudes
            t = self.data.reshape((2,) * self.nq)
            rest = [q for q in range(self.nq) if q not in keep]
            t = np.transpose(t, keep + rest).reshape(1 << len(keep), -1)
            return FakeState(t @ t.conj().T, DM, len(keep))
        if not keep:
            return FakeState(np.trace(self.data).reshape(1, 1), DM, 0)
        assert keep == sorted(keep)
        a, _ = orc.ptrace_arbitrary(self.data, self.nq, keep) if len(keep) < self.nq else (self.data, None)
        return FakeState(a, DM, len(keep))

    @staticmethod
    def scatter_product(a, b, a_positions, b_positions=(), scale=None):
        bd = b.data if b is not None else np.array([], dtype=complex)
        out = orc.interweave(a.data, bd, list(a_positions)) if bd.size else a.data.copy()
        if b is not None and b.nq == 0 and bd.size:
            pass  # interweave already multiplied by the 1x1 factor
        if scale is not None:
            out = out * scale
        return FakeState(out, DM, a.nq + (b.nq if b is not None else 0))

    def probs(self, qubits):
        if self.kind == KET:
            return orc.ket_probs(self.data, self.nq, list(qubits))
        d = np.diag(self.data).reshape((2,) * self.nq)
        others = tuple(q for q in range(self.nq) if q not in qubits)
        d = d.sum(axis=others) if others else d
        asc = sorted(qubits)
        return np.abs(np.transpose(d, [asc.index(q) for q in qubits]).reshape(-1))

    def probs_basis(self, qubits, basis_kets):
        """the reference's outcome loop (measurement.py:147-155) on the listed qubits of this state"""
        kets = [np.asarray(k, dtype=complex).reshape(-1) for k in basis_kets]
        dens = [np.outer(k, k) for k in kets]
        b = orc.ilog2(kets[0].shape[0])
        qubits = list(qubits)
        if self.kind == KET and self.nq > 12:
            # psi psi^dagger of a large ket does not fit: the rotated-amplitude form (equal for real bases, and what the
            # device computes for a ket-mode register)
            assert qubits == sorted(qubits)
            return orc.basis_weights(self.data, self.nq, qubits, kets)
        rho = self.data if self.kind == DM else np.outer(self.data, self.data.conj())
        if qubits != list(range(self.nq)):
            assert qubits == sorted(qubits)
            rho, _ = orc.ptrace_arbitrary(rho, self.nq, qubits)
        f = len(qubits) // b
        return np.array([abs(np.trace(np.matmul(rho, orc.basis_projector(f, i, dens)[0]))) for i in range(len(dens) ** f)])

    def apply_gate_rc(self, row_matrix, col_matrix, first_target=0):
        assert self.kind == DM
        ref = row_matrix if row_matrix is not None else col_matrix
        k = orc.ilog2(np.asarray(ref).shape[0])
        if row_matrix is not None:
            self.data = orc.embed_gate(self.nq, first_target, np.asarray(row_matrix, dtype=complex)) @ self.data
        if col_matrix is not None:
            self.data = self.data @ orc.embed_gate(self.nq, first_target, np.asarray(col_matrix, dtype=complex)).T
        return self

    @classmethod
    def diagonal(cls, weights, device=0):
        w = np.asarray(weights, dtype=float)
        return cls(np.diag(w).astype(complex), DM, orc.ilog2(w.shape[0]))

    def as_density(self):
        return self if self.kind == DM else self.outer(True)

    def outer(self, conj=True):
        return FakeState(np.outer(self.data, self.data.conj() if conj else self.data), DM, self.nq)
