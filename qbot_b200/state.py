"""Device-resident register: Python handle over the C ABI (include/qbot_b200.h).

``DeviceState`` is what ``localNameSpace['state']`` holds when the backend is installed.  It
honours the contract the reference's ops and user expressions rely on (SURVEY.md F11,
8(b) "State contract"): ``.shape / .size / .ndim / .dtype``, conversion with ``np.asarray``,
indexing and arithmetic (through a host copy made on demand).  Qubit arguments use the
reference's numbering (qubit 0 = most significant index bit, qbot/qgates.py:161-182); the
conversion to index bits happens here.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional, Sequence

import numpy as np

from . import _lib

KET, DM = 0, 1
HOST_MATERIALIZE_LIMIT = 1 << 31      # bytes; larger states refuse implicit np.asarray()


_ONE_BIT = [(C.c_int * 1)(b) for b in range(64)]
_fn_cache = {}


def _apply_gate_fn():
    f = _fn_cache.get('qb_apply_gate')
    if f is None:
        f = _fn_cache['qb_apply_gate'] = _lib.load().qb_apply_gate
    return f


def _cptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _cmat(m, dim=None) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(m), dtype=np.complex128)
    if dim is not None and a.shape != (dim, dim):
        raise ValueError(f"matrix must be {dim}x{dim}")
    return a


def marshal_batched(nq: int, nbranch: int, matrices, target_qubits, controls=None, enable=None):
    """Host-side preparation of one batched launch (qb_apply_gate_batched): [nbranch, 2^k, 2^k] complex128
    matrices, int32 target index bits [nbranch * k] (qubit q <-> bit nq-1-q, matrix msb first), uint64 control
    masks, uint8 enable flags or None.  Vectorised: at 4096 branches the per-branch Python loops this replaces
    cost more than the kernel (1.5 ms against 1.36 ms)."""
    m = np.ascontiguousarray(np.asarray(matrices), dtype=np.complex128)
    b = nbranch
    if m.ndim != 3 or m.shape[0] != b or m.shape[1] != m.shape[2]:
        raise ValueError("matrices must be [nbranch, 2^k, 2^k]")
    k = int(m.shape[1]).bit_length() - 1
    try:
        tq = np.asarray(target_qubits, dtype=np.int64)
    except (ValueError, TypeError):
        tq = None                                    # ragged
    if tq is None or tq.ndim != 2 or tq.shape[1] != k:
        raise ValueError("every branch needs k target qubits")
    if tq.shape[0] != b:
        raise ValueError("one target list per branch")
    tb = np.ascontiguousarray(nq - 1 - tq, dtype=np.int32).reshape(-1)
    masks = np.zeros(max(b, 1), dtype=np.uint64)
    if controls is not None:
        cs = None
        try:
            cs = np.asarray(controls, dtype=np.int64)
        except (ValueError, TypeError):
            pass
        if cs is not None and cs.ndim == 2 and cs.shape[0] == b:
            if cs.shape[1]:
                masks[:b] = np.bitwise_or.reduce(np.uint64(1) << (nq - 1 - cs).astype(np.uint64), axis=1)
        else:                                        # ragged control lists
            for i, c in enumerate(controls):
                cm = 0
                for q in c:
                    cm |= 1 << (nq - 1 - int(q))
                masks[i] = cm
    en = None if enable is None else np.ascontiguousarray(np.asarray(enable, dtype=np.uint8))
    return m, k, tb, masks, en


class DeviceState:
    """A ket (n qubits -> 2^n amplitudes) or density matrix (2^n x 2^n) in HBM, optionally
    with a leading branch axis (``nbranch`` independent copies: the ProbVal batch)."""
    _qb_device_state = True
    __array_priority__ = 1000
    _shared = True        # ownership flag of the DSL ops (host/ops.py _exclusively_owned): conservative default
    _version = 0          # bumped by every in-place update: lazily computed views of the register (host/ops.py _LazyReducedDensity) check it

    def __init__(self, handle, kind: int, nq: int, nbranch: int = 1):
        self._h = C.c_void_p(handle) if not isinstance(handle, C.c_void_p) else handle
        self.kind = kind
        self.nq = nq
        self.nbranch = nbranch
        self._host_cache = None
        self._keepalive = None

    # ---- construction ---------------------------------------------------------------------------
    @classmethod
    def create(cls, kind: int, nq: int, nbranch: int = 1, device: int = 0) -> "DeviceState":
        h = C.c_void_p()
        _lib.call('qb_create', C.byref(h), kind, nq, nbranch, device)
        return cls(h, kind, nq, nbranch)

    @classmethod
    def zero_state(cls, nq: int, kind: int = KET, nbranch: int = 1, device: int = 0) -> "DeviceState":
        s = cls.create(kind, nq, nbranch, device)
        _lib.call('qb_init_basis', s._h, 0)
        return s

    @classmethod
    def from_host(cls, array, device: int = 0) -> "DeviceState":
        """2-D square complex array -> density matrix; 1-D -> ket."""
        a = np.ascontiguousarray(np.asarray(array), dtype=np.complex128)
        if a.ndim == 2:
            if a.shape[0] != a.shape[1] or a.shape[0] & (a.shape[0] - 1):
                raise ValueError("density matrix must be 2^n x 2^n")
            kind, nq = DM, int(a.shape[0]).bit_length() - 1
        elif a.ndim == 1:
            if a.shape[0] & (a.shape[0] - 1):
                raise ValueError("ket length must be a power of two")
            kind, nq = KET, int(a.shape[0]).bit_length() - 1
        else:
            raise ValueError("state must be 1-D or 2-D")
        s = cls.create(kind, nq, 1, device)
        _lib.call('qb_upload', s._h, _cptr(a), a.nbytes)
        return s

    @classmethod
    def from_kets(cls, kets: np.ndarray, device: int = 0) -> "DeviceState":
        """[B, 2^n] array -> batched ket state."""
        a = np.ascontiguousarray(np.asarray(kets), dtype=np.complex128)
        nq = int(a.shape[1]).bit_length() - 1
        s = cls.create(KET, nq, a.shape[0], device)
        _lib.call('qb_upload', s._h, _cptr(a), a.nbytes)
        return s

    @classmethod
    def product(cls, factors: Sequence[np.ndarray], kind: int = KET, device: int = 0) -> "DeviceState":
        """Product state from per-qubit 2-vectors (ket) or 2x2 matrices (density), qubit 0 first."""
        per = 2 if kind == KET else 4
        v = np.ascontiguousarray(np.stack([np.asarray(f, dtype=np.complex128).reshape(per) for f in factors]))
        s = cls.create(kind, len(factors), 1, device)
        _lib.call('qb_init_product', s._h, _cptr(v), 0)
        return s

    @classmethod
    def product_batch(cls, factors: np.ndarray, device: int = 0) -> "DeviceState":
        """[B, n, 2] per-branch per-qubit 2-vectors -> batched product kets."""
        v = np.ascontiguousarray(np.asarray(factors, dtype=np.complex128))
        b, nq = v.shape[0], v.shape[1]
        s = cls.create(KET, nq, b, device)
        _lib.call('qb_init_product', s._h, _cptr(v), 1)
        return s

    def __del__(self):
        h = getattr(self, '_h', None)
        if h is not None and h.value:
            try:
                _lib.load().qb_destroy(h)
            except Exception:
                pass
            self._h = None

    def clone(self) -> "DeviceState":
        h = C.c_void_p()
        _lib.call('qb_clone', self._h, C.byref(h))
        return DeviceState(h, self.kind, self.nq, self.nbranch)

    # ---- ndarray contract -----------------------------------------------------------------------
    @property
    def dim(self) -> int:
        return 1 << self.nq

    @property
    def shape(self):
        base = (self.dim, self.dim) if self.kind == DM else (self.dim,)
        return base if self.nbranch == 1 else (self.nbranch,) + base

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        n = 1
        for d in self.shape:
            n *= d
        return n

    @property
    def dtype(self):
        return np.dtype(np.complex128)

    @property
    def nbytes(self):
        return self.size * 16

    def _dirty(self):
        self._host_cache = None
        self._version += 1

    def to_host(self) -> np.ndarray:
        if self._host_cache is None:
            if self.nbytes > HOST_MATERIALIZE_LIMIT:
                raise MemoryError(f"state of {self.nbytes} bytes is too large to copy to the host implicitly; "
                                  "use download_range()")
            out = np.empty(self.shape, dtype=np.complex128)
            _lib.call('qb_download', self._h, _cptr(out), out.nbytes)
            self._host_cache = out
        return self._host_cache

    def download_range(self, first: int, count: int) -> np.ndarray:
        out = np.empty(count, dtype=np.complex128)
        _lib.call('qb_download_range', self._h, first, count, _cptr(out))
        return out

    def __array__(self, dtype=None, copy=None):
        a = self.to_host()
        if dtype is not None and np.dtype(dtype) != a.dtype:
            return a.astype(dtype)
        return a.copy() if copy else a

    def __getattr__(self, name):
        # anything else a user expression asks of `state` is answered by the host copy
        if name.startswith('_'):
            raise AttributeError(name)
        return getattr(self.to_host(), name)

    def __getitem__(self, idx):
        return self.to_host()[idx]

    def __len__(self):
        return self.shape[0]

    def __iter__(self):
        return iter(self.to_host())

    def __repr__(self):
        k = 'dm' if self.kind == DM else 'ket'
        return f"DeviceState({k}, qubits={self.nq}, branches={self.nbranch})"

    def __str__(self):
        return str(self.to_host())

    # ---- gates (qgates.genGateForFullHilbertSpace / genMultiControlledGate / applyGate) ---------
    def _bit(self, qubit: int) -> int:
        return self.nq - 1 - int(qubit)

    def apply_gate(self, matrix, first_target: int = 0, controls: Iterable[int] = ()) -> "DeviceState":
        m = matrix if type(matrix) is np.ndarray and matrix.dtype == np.complex128 else _cmat(matrix)
        dim = m.shape[0]
        k = dim.bit_length() - 1
        if m.ndim != 2 or m.shape[1] != dim or dim != 1 << k:
            raise Exception("gate size must be power of 2")
        nq = self.nq
        if first_target < 0 or first_target + k - 1 >= nq:
            raise IndexError(f"{k} qubit gate does not fit the {nq} qubit hilbertspace when started on qubit {first_target}")
        # (this is the per-line cost of a `gate` op: the matrix travels as a bytes object -- the library copies it
        # while queueing -- and one-target bit lists are shared constants; ndarray.ctypes alone costs 2-4 us)
        bits = _ONE_BIT[nq - 1 - first_target] if k == 1 else _lib.int_array([nq - 1 - first_target - j for j in range(k)])
        cmask = 0
        for c in controls:
            cmask |= 1 << (nq - 1 - int(c))
        rc = _apply_gate_fn()(self._h, m.tobytes(), k, bits, cmask)
        if rc:
            _lib.check(rc)
        self._host_cache = None
        self._version += 1
        return self

    def apply_gate_bits(self, matrix, target_bits: Sequence[int], control_mask: int = 0) -> "DeviceState":
        """Gate on arbitrary (non-contiguous) index bits; target_bits[0] = matrix MSB."""
        m = _cmat(matrix)
        k = len(target_bits)
        _lib.call('qb_apply_gate', self._h, _cptr(m), k, _lib.int_array(target_bits), control_mask)
        self._dirty()
        return self

    @staticmethod
    def pack_circuit(nq: int, items) -> tuple:
        """items: [(matrix, first_target, controls)] in reference qubit numbering -> the argument
        block of qb_apply_gates (build it once for a circuit that is applied repeatedly)."""
        n = len(items)
        ks = (C.c_int * max(n, 1))()
        tbs = (C.c_int * (14 * max(n, 1)))()
        cms = (C.c_uint64 * max(n, 1))()
        mats = []
        for i, (m, t, controls) in enumerate(items):
            m = np.ascontiguousarray(np.asarray(m, dtype=np.complex128))
            k = int(m.shape[0]).bit_length() - 1
            if t < 0 or t + k - 1 >= nq:
                raise IndexError(f"{k} qubit gate does not fit the {nq} qubit hilbertspace when started on qubit {t}")
            ks[i] = k
            for j in range(k):
                tbs[14 * i + j] = nq - 1 - (t + j)
            cm = 0
            for c in controls:
                cm |= 1 << (nq - 1 - int(c))
            cms[i] = cm
            mats.append(m.reshape(-1))
        allm = np.ascontiguousarray(np.concatenate(mats)) if mats else np.zeros(1, dtype=np.complex128)
        return n, ks, tbs, cms, allm

    def apply_circuit(self, packed) -> "DeviceState":
        """Queue a whole packed circuit with ONE call into the library (qb_apply_gates)."""
        n, ks, tbs, cms, allm = packed
        _lib.call('qb_apply_gates', self._h, n, ks, tbs, cms, _cptr(allm))
        self._dirty()
        return self

    def apply_swap(self, qubit_a: int, qubit_b: int) -> "DeviceState":
        _lib.call('qb_apply_swap', self._h, self._bit(qubit_a), self._bit(qubit_b))
        self._dirty()
        return self

    def apply_gate_batched(self, matrices, first_targets: Sequence[int], controls: Sequence[Iterable[int]] = None,
                           enable: Sequence[bool] = None) -> "DeviceState":
        """Branch b applies matrices[b] on qubits first_targets[b].. with controls[b] (one launch)."""
        m = np.asarray(matrices)
        k = int(m.shape[1]).bit_length() - 1
        ft = np.asarray(first_targets, dtype=np.int64).reshape(-1)
        bad = (ft < 0) | (ft + k - 1 >= self.nq)
        if bad.any():
            t = int(ft[bad][0])
            raise IndexError(f"{k} qubit gate does not fit the {self.nq} qubit hilbertspace when started on qubit {t}")
        return self.apply_gate_batched_qubits(m, ft[:, None] + np.arange(k, dtype=np.int64)[None, :], controls, enable)

    def apply_gate_batched_qubits(self, matrices, target_qubits: Sequence[Sequence[int]],
                                  controls: Sequence[Iterable[int]] = None, enable: Sequence[bool] = None) -> "DeviceState":
        """Same, with an explicit (possibly non-contiguous) qubit list per branch;
        target_qubits[b][0] carries the matrix's most significant index bit."""
        m, k, tb, masks, en_arr = marshal_batched(self.nq, self.nbranch, matrices, target_qubits, controls, enable)
        _lib.call('qb_apply_gate_batched', self._h, _cptr(m), k, tb.ctypes.data_as(C.POINTER(C.c_int)),
                  masks.ctypes.data_as(C.POINTER(C.c_uint64)), None if en_arr is None else _cptr(en_arr))
        self._dirty()
        return self

    def branch_view(self, b: int) -> "DeviceState":
        """Non-owning single-branch handle on branch b of a batched state."""
        p = C.c_void_p()
        _lib.call('qb_device_ptr', self._h, C.byref(p))
        per = (1 << (self.nq if self.kind == KET else 2 * self.nq)) * 16
        h = C.c_void_p()
        _lib.call('qb_create_external', C.byref(h), self.kind, self.nq, 1, self.device, C.c_void_p(p.value + b * per), None)
        v = DeviceState(h, self.kind, self.nq, 1)
        v._keepalive = self
        return v

    def apply_branch_gates(self, items) -> "DeviceState":
        """items[b] = (matrix, target qubits, controls) or None; gates of different sizes per
        branch (rare) are applied branch by branch through views."""
        for b, it in enumerate(items):
            if it is None:
                continue
            m, qubits, controls = it
            v = self.branch_view(b)
            cmask = 0
            for c in controls:
                cmask |= 1 << self._bit(c)
            v.apply_gate_bits(m, [self._bit(q) for q in qubits], cmask)
            v.sync()
        self._dirty()
        return self

    def flush(self):
        _lib.call('qb_flush', self._h)

    def sync(self):
        _lib.call('qb_sync', self._h)

    def set_fusion(self, on: bool):
        _lib.call('qb_set_fusion', self._h, 1 if on else 0)

    def set_jit(self, mode: int):
        """Sweep specialisation: 0 never, 1 when a plan repeats on a large state (default), 2 always."""
        _lib.call('qb_set_jit', self._h, int(mode))

    # ---- measurement ------------------------------------------------------------------------------
    def probs(self, qubits: Sequence[int]) -> np.ndarray:
        """Computational-basis outcome weights of the listed qubits (first listed = most
        significant outcome bit).  Shape [2^m] or [nbranch, 2^m]."""
        m = len(qubits)
        out = np.empty((self.nbranch, 1 << m), dtype=np.float64)
        _lib.call('qb_probs', self._h, _lib.int_array([self._bit(q) for q in qubits]), m, _cptr(out))
        return out[0] if self.nbranch == 1 else out

    def probs_basis(self, qubits: Sequence[int], basis_kets) -> np.ndarray:
        """Outcome weights of the listed qubits in a product of measurement bases
        (measurement.permuteBasis + the outcome loop, qbot/measurement.py:88-101, 147-155): the
        qubits are taken in groups of b = log2(len(kets[0])) as listed, outcome index = base-2^b number
        whose most significant digit belongs to the first group.  Runs on the device (two gate sweeps on
        a scratch copy + the binned read); the register is not modified."""
        w = np.ascontiguousarray(np.stack([np.asarray(k, dtype=np.complex128).reshape(-1) for k in basis_kets]))
        b = int(w.shape[1]).bit_length() - 1
        if w.shape != (1 << b, 1 << b):
            raise ValueError("a measurement basis needs 2^b kets of 2^b amplitudes")
        m = len(qubits)
        out = np.empty((self.nbranch, 1 << m), dtype=np.float64)
        _lib.call('qb_probs_basis', self._h, _lib.int_array([self._bit(q) for q in qubits]), m, _cptr(w), b, _cptr(out))
        return out[0] if self.nbranch == 1 else out

    def apply_gate_rc(self, row_matrix, col_matrix, first_target: int = 0) -> "DeviceState":
        """Density matrices: rho <- R rho C^T on the contiguous qubits starting at first_target
        (either matrix may be None).  The non-conjugating forms of the reference live here, see
        qb_apply_gate_rc in the header."""
        r = None if row_matrix is None else _cmat(row_matrix)
        c = None if col_matrix is None else _cmat(col_matrix)
        ref = r if r is not None else c
        k = int(ref.shape[0]).bit_length() - 1
        bits = _lib.int_array([self._bit(first_target + j) for j in range(k)])
        _lib.call('qb_apply_gate_rc', self._h, None if r is None else _cptr(r), None if c is None else _cptr(c), k, bits)
        self._dirty()
        return self

    @classmethod
    def diagonal(cls, weights: Sequence[float], device: int = 0) -> "DeviceState":
        """Density matrix diag(weights), built on the device."""
        w = np.ascontiguousarray(np.asarray(weights, dtype=np.float64).reshape(-1))
        nq = int(w.shape[0]).bit_length() - 1
        if w.shape[0] != 1 << nq:
            raise ValueError("diagonal: needs 2^n weights")
        s = cls.create(DM, nq, 1, device)
        _lib.call('qb_init_diag', s._h, _cptr(w))
        return s

    def norm2(self) -> np.ndarray:
        out = np.empty(self.nbranch, dtype=np.float64)
        _lib.call('qb_norm2', self._h, _cptr(out))
        return out

    def project_renorm(self, qubits: Sequence[int], outcome: int) -> "DeviceState":
        _lib.call('qb_project_renorm', self._h, _lib.int_array([self._bit(q) for q in qubits]), len(qubits), outcome)
        self._dirty()
        return self

    # ---- density structure (density.partialTraceArbitrary / interweaveDensities / ensemble) ----
    def ptrace_keep(self, keep_qubits: Sequence[int]) -> "DeviceState":
        """Trace out everything except ``keep_qubits``; the result's qubit i is keep_qubits[i]."""
        h = C.c_void_p()
        bits = _lib.int_array([self._bit(q) for q in keep_qubits])
        _lib.call('qb_ptrace', self._h, bits, len(keep_qubits), C.byref(h))
        return DeviceState(h, DM, len(keep_qubits), 1)

    @staticmethod
    def scatter_product(a: "DeviceState", b: Optional["DeviceState"], a_positions: Sequence[int],
                        b_positions: Sequence[int] = (), scale: complex = None) -> "DeviceState":
        """rho_A (x) rho_B with A's qubit i at final qubit a_positions[i], B's at b_positions[i]."""
        n = a.nq + (b.nq if b is not None else 0)
        # the C entry point takes no lengths: it reads a.nq + b.nq positions
        if len(a_positions) != a.nq or len(b_positions) != (b.nq if b is not None else 0):
            raise ValueError(f"scatter_product: {len(a_positions)} + {len(b_positions)} positions for a {a.nq} + "
                             f"{b.nq if b is not None else 0} qubit product")
        abits = _lib.int_array([n - 1 - p for p in a_positions])
        bbits = _lib.int_array([n - 1 - p for p in b_positions])
        h = C.c_void_p()
        sc = None
        if scale is not None:
            sc_arr = np.array([complex(scale).real, complex(scale).imag], dtype=np.float64)
            sc = _cptr(sc_arr)
        _lib.call('qb_scatter_product', a._h, b._h if b is not None else None, abits, bbits, sc, C.byref(h))
        return DeviceState(h, DM, n, 1)

    @staticmethod
    def mix(probs: Sequence[float], states: Sequence["DeviceState"]) -> "DeviceState":
        if len(probs) != len(states):
            raise Exception("number of state vectors an number of probabilites must equal")
        s0 = states[0]
        for s in states:
            if s.shape != s0.shape or s.kind != s0.kind:
                raise ValueError(f"operands could not be broadcast together with shapes {s0.shape} {s.shape}")
        if s0.kind == KET:
            # an ensemble of kets is the density matrix sum_i p_i psi_i psi_i^T (ProbVal.toDensityMatrix,
            # qbot/probVal.py:99-111 -- np.outer without conjugation), never sum_i p_i psi_i
            if s0.nq > 13:
                raise ValueError(f"an ensemble of {s0.nq}-qubit kets needs a 4^{s0.nq}-entry density matrix")
            states = [s.outer(False) for s in states]
            s0 = states[0]
        arr = (C.c_void_p * len(states))(*[s._h for s in states])
        p = (C.c_double * len(probs))(*[float(x) for x in probs])
        h = C.c_void_p()
        _lib.call('qb_mix', arr, p, len(states), C.byref(h))
        return DeviceState(h, s0.kind, s0.nq, s0.nbranch)

    def mix_branches(self, probs: Sequence[float]) -> "DeviceState":
        p = np.ascontiguousarray(np.asarray(probs, dtype=np.float64).reshape(-1))
        if p.size != self.nbranch:
            raise ValueError("one weight per branch")
        h = C.c_void_p()
        _lib.call('qb_mix_branches', self._h, p.ctypes.data_as(C.POINTER(C.c_double)), C.byref(h))
        return DeviceState(h, self.kind, self.nq, 1)

    def broadcast(self, nbranch: int) -> "DeviceState":
        dst = DeviceState.create(self.kind, self.nq, nbranch, self.device)
        _lib.call('qb_broadcast', self._h, dst._h)
        return dst

    def as_density(self) -> "DeviceState":
        """The density matrix of this state (a ket becomes psi psi^dagger)."""
        return self if self.kind == DM else self.outer(True)

    def outer(self, conj: bool = True) -> "DeviceState":
        h = C.c_void_p()
        _lib.call('qb_outer', self._h, 1 if conj else 0, C.byref(h))
        return DeviceState(h, DM, self.nq, 1)

    @property
    def device(self) -> int:
        d = C.c_int()
        _lib.call('qb_info', self._h, None, None, None, C.byref(d))
        return d.value

    # ---- instrumentation --------------------------------------------------------------------------
    def stats(self) -> dict:
        st = _lib.QbStats()
        _lib.call('qb_get_stats', self._h, C.byref(st))
        return {k: int(getattr(st, k)) for k, _ in st._fields_}

    def reset_stats(self):
        _lib.call('qb_reset_stats', self._h)

    def timer_start(self):
        _lib.call('qb_timer_start', self._h)

    def timer_stop(self) -> float:
        ms = C.c_float()
        _lib.call('qb_timer_stop', self._h, C.byref(ms))
        return ms.value


def _binop(name):
    def f(self, other):
        return getattr(self.to_host(), name)(np.asarray(other) if isinstance(other, DeviceState) else other)
    f.__name__ = name
    return f


for _n in ('__add__', '__radd__', '__sub__', '__rsub__', '__mul__', '__rmul__', '__truediv__', '__rtruediv__',
           '__matmul__', '__rmatmul__', '__eq__', '__ne__', '__lt__', '__le__', '__gt__', '__ge__', '__pow__'):
    setattr(DeviceState, _n, _binop(_n))
DeviceState.__hash__ = None
DeviceState.__neg__ = lambda self: -self.to_host()
DeviceState.__abs__ = lambda self: abs(self.to_host())
