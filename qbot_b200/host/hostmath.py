"""Small host-side matrices of the DSL namespace: gate constructors, bases, tensor helpers.

These are the 2^k x 2^k *inputs* of the state path (k = a few qubits), not the path itself;
they stay numpy on the host exactly as in the reference (qbot/qgates.py:18-74, 136-158;
qbot/density.py:7-74; qbot/measurement.py:72-86; qbot/basis.py).  Values are computed with
the same floating-point expressions so that programs see identical constants.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple, Union

import numpy as np


def ilog2(x: int) -> int:
    return 0 if x == 0 else int(x).bit_length() - 1


def ensure_square(a) -> int:
    if a.size == 0:
        return 0
    if a.ndim != 2:
        raise Exception("array must be 2 dimensional")
    if a.shape[0] != a.shape[1]:
        raise Exception("array must be square")
    return a.shape[0]


# ---- gate constructors ------------------------------------------------------------------------
def x_rot(theta):
    s, c = np.sin(theta / 2), np.cos(theta / 2)
    return np.array([[c, -1j * s], [-1j * s, c]], dtype=complex)


def y_rot(theta):
    s, c = np.sin(theta / 2), np.cos(theta / 2)
    return np.array([[c, -s], [s, c]], dtype=complex)


def z_rot(theta):
    return np.array([[np.exp(-1j * theta / 2), 0], [0, np.exp(1j * theta / 2)]], dtype=complex)


def qft(num_qubits: int):
    assert isinstance(num_qubits, int)
    size = 2 ** num_qubits
    roots = np.exp(2j * np.pi / size * np.arange(size)) / np.sqrt(size)
    k = np.arange(size)
    return roots[np.outer(k, k) % size].astype(complex)


def simons(num_qubits: int, f: Callable):
    """U_f |x>|b> = |x>|b xor f(x)> on num_qubits qubits (last qubit is b)."""
    size = 2 ** num_qubits
    u = np.zeros((size, size), dtype=complex)
    for i in range(size):
        x, b = i >> 1, i & 1
        u[i][(x << 1) + (f(x) + b) % 2] = 1
    return u


def permutation(dim: int, state_map: Callable[[int], int]):
    u = np.zeros((dim, dim), dtype=complex)
    for i in range(dim):
        u[state_map(i)][i] = 1
    return u


def swap_matrix(num_qubits: int, a: int, b: int):
    """Qubit transposition as a matrix.  (The reference's genSwapGate, qgates.py:77-133, equals
    this for n <= 4 and is not a transposition for n >= 5 -- SURVEY.md F5; the definition is
    what is implemented here.)"""
    if a == b:
        return np.eye(2 ** num_qubits)
    if max(a, b) >= num_qubits:
        raise Exception("getSwapGate Requires numQubits > q2 and q1")
    ba, bb = num_qubits - 1 - a, num_qubits - 1 - b

    def sm(i):
        x = ((i >> ba) ^ (i >> bb)) & 1
        return i ^ ((x << ba) | (x << bb))

    return permutation(2 ** num_qubits, sm)


def shift_matrix(num_qubits: int, up: bool = True, shifts: int = 1):
    dim = 2 ** num_qubits
    if up:
        return permutation(dim, lambda s: ((s << shifts) % dim) | ((s << shifts) // dim))
    return permutation(dim, lambda s: (s >> shifts) | ((s & (2 ** shifts - 1)) << (num_qubits - shifts)))


# ---- tensor helpers -----------------------------------------------------------------------------
LAZY_MIN_QUBITS = 14      # smaller products are built on the host exactly as the reference does
LAZY_MIN_QUBITS_DM = 8    # ... density matrices from 8 qubits on (1 MiB): the 12-qubit |0..0><0..0| of BASELINE config 3
                          # costs ~0.25 s as a host kron chain + 256 MiB upload and ~0.05 ms as a device fill


class LazyProduct:
    """kron(factors[0], factors[1], ...) kept as a descriptor.

    The reference builds initial states with kron chains on the host (density.py:7-29), which
    stops being possible around 14 qubits (a 2^n x 2^n complex128 array).  From LAZY_MIN_QUBITS
    on, tensorProd / tensorExp return this ndarray-like object instead; `qset` hands its
    factors to the device-side constructor (qb_init_product) and no 2^n array ever exists on
    the host.  All factors are kets (1-D) or all are density matrices (2-D)."""
    _qb_lazy_product = True

    def __init__(self, factors):
        flat = []
        for f in factors:
            if getattr(f, '_qb_lazy_product', False):
                flat.extend(f.factors)
            else:
                flat.append(np.asarray(f, dtype=complex))
        nd = {f.ndim for f in flat}
        if len(nd) != 1 or next(iter(nd)) not in (1, 2):
            raise ValueError("tensor product of kets with density matrices")
        self.factors = flat
        self.ndim = flat[0].ndim
        self.num_qubits = sum(ilog2(f.shape[0]) for f in flat)
        self.dtype = np.dtype(complex)

    @property
    def shape(self):
        d = 1 << self.num_qubits
        return (d,) if self.ndim == 1 else (d, d)

    @property
    def size(self):
        return (1 << self.num_qubits) ** self.ndim

    def single_qubit_factors(self):
        """per-qubit 2-vectors / 2x2 matrices, qubit 0 first, or None if a factor spans several qubits"""
        if all(f.shape[0] == 2 for f in self.factors):
            return self.factors
        return None

    def materialize(self):
        out = self.factors[0]
        for f in self.factors[1:]:
            out = np.kron(out, f)
        return out

    def __array__(self, dtype=None, copy=None):
        a = self.materialize()
        return a if dtype is None else a.astype(dtype)

    def __repr__(self):
        return f"LazyProduct({len(self.factors)} factors, {self.num_qubits} qubits, {'ket' if self.ndim == 1 else 'density'})"

    def __getattr__(self, name):
        # anything else a user expression asks of the product is answered by the array it stands for
        if name.startswith('_'):
            raise AttributeError(name)
        return getattr(self.materialize(), name)

    def __getitem__(self, idx):
        return self.materialize()[idx]

    def __len__(self):
        return self.shape[0]

    def __iter__(self):
        return iter(self.materialize())

    __hash__ = None
    __array_priority__ = 1000


def _lazy_binop(name):
    def f(self, other):
        return getattr(self.materialize(), name)(np.asarray(other) if is_lazy(other) else other)
    f.__name__ = name
    return f


for _n in ('__add__', '__radd__', '__sub__', '__rsub__', '__mul__', '__rmul__', '__truediv__', '__rtruediv__',
           '__matmul__', '__rmatmul__', '__eq__', '__ne__', '__pow__', '__neg__', '__abs__'):
    setattr(LazyProduct, _n, _lazy_binop(_n) if _n not in ('__neg__', '__abs__') else
            (lambda nm: (lambda self: getattr(self.materialize(), nm)()))(_n))


def is_lazy(x) -> bool:
    return getattr(x, '_qb_lazy_product', False)


def tensor_prod(*parts):
    parts = [p for p in parts if p.size != 0]
    if not parts:
        return np.array([], dtype=complex)
    if len({p.ndim for p in parts}) == 1 and parts[0].ndim in (1, 2) and \
            sum(ilog2(p.shape[0]) for p in parts) >= (LAZY_MIN_QUBITS if parts[0].ndim == 1 else LAZY_MIN_QUBITS_DM):
        return LazyProduct(parts)
    out = None
    for p in parts:
        p = np.asarray(p)
        out = p if out is None else np.kron(out, p)
    return out


def tensor_exp(state, n: int):
    if n == 0:
        return np.eye(state.shape[0], dtype=complex)
    return tensor_prod(*([state] * n))


def ket_to_density(ket):
    return np.outer(ket, ket)          # no conjugation: SURVEY.md F2


def kets_to_density(kets, probs=None):
    if probs is None:
        return ket_to_density(kets[0])
    if len(kets) != len(probs):
        raise Exception("number of state vectors an number of probabilites must equal")
    acc = np.zeros((kets[0].shape[0],) * 2, dtype=complex)
    for p, k in zip(probs, kets):
        acc += p * np.outer(k, k)
    return acc


def kets_to_density_zipped(pairs):
    if len(pairs) == 0:
        return np.array([], dtype=complex)
    if len(pairs) == 1:
        return ket_to_density(pairs[0][1])
    acc = np.zeros((pairs[0][1].shape[0],) * 2, dtype=complex)
    for p, k in pairs:
        acc += p * np.outer(k, k)
    return acc


def density_to_kets(rho):
    """Eigen-decomposition into (weight, vector) pairs, as the reference's densityToKets
    (density.py:232-240) including its row-indexing of the eigenvector matrix."""
    ensure_square(rho)
    vals, vecs = np.linalg.eig(rho)
    return [(abs(v), vecs[i]) for i, v in enumerate(vals) if v != 0]


class Basis:
    __slots__ = ('names', 'density', 'kets', 'numQubits', 'ketSymbols', 'gateSymbol')

    def __init__(self, names, kets, ketSymbols, gateSymbol):
        if len(ketSymbols) != len(kets):
            raise Exception("basis must have same number of ketSymbols and kets")
        self.names = names
        self.kets = kets
        self.numQubits = ilog2(kets[0].shape[0])
        self.ketSymbols = ketSymbols
        self.gateSymbol = gateSymbol
        self.density = [kets_to_density([k]) for k in kets]

    def __getitem__(self, i):
        return self.density[i]


_s = 2 ** (-1 / 2)
computation = Basis(['comp', 'computation', 'computational', 'compBasis', 'computationBasis', 'computationalBasis'],
                    [np.array([1, 0], dtype=complex), np.array([0, 1], dtype=complex)],
                    ["|0〉", "|1〉"], '∡')
hadamard = Basis(['hadamard', 'had', 'hada', 'hadamardBasis', 'hadBasis', 'hadaBasis'],
                 [_s * np.array([1, 1], dtype=complex), _s * np.array([1, -1], dtype=complex)],
                 ["|+〉", "|-〉"], '∡ ±')
bell = Basis(['bell', 'epr', 'bellBasis', 'eprBasis'],
             [_s * np.array([1, 0, 0, 1], dtype=complex), _s * np.array([0, 1, 1, 0], dtype=complex),
              _s * np.array([1, 0, 0, -1], dtype=complex), _s * np.array([0, 1, -1, 0], dtype=complex)],
             ["|β₀₀〉", "|β₀₁〉", "|β₁₀〉", "|β₁₁〉"], '∡ β')
all_bases = [computation, hadamard, bell]


def tensor_permute(num_factors: int, n: int, d: Union[Sequence[np.ndarray], Basis]):
    """n-th tensor permutation of the list's matrices, least significant digit rightmost."""
    if isinstance(d, Basis):
        d = d.density
    out = np.array([], dtype=complex)
    rem = n
    for _ in range(num_factors):
        out = tensor_prod(d[rem % len(d)], out)
        rem //= len(d)
    return out


def permute_basis(num_factors: int, n: int, basis: Basis) -> Tuple[np.ndarray, str]:
    out = np.array([], dtype=complex)
    sym = ''
    rem = n
    for _ in range(num_factors):
        k = rem % len(basis.density)
        out = tensor_prod(basis.density[k], out)
        sym = basis.ketSymbols[k] + sym
        rem //= len(basis.density)
    return out, sym
