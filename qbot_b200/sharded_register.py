"""A ket sharded over the GPUs of one box as the DSL's register (BASELINE config 5).

The reference keeps ONE register, ``localNameSpace['state']`` (qbot/interpreter.py:218-224), and six
ops drive it (qbot/operators.py:133-188, 255-329, 364-428).  When a program is run by one process
per GPU (``torchrun``; every rank executes the same program text) and a ket of at least
``min_qubits`` qubits is put into the register --

    qset tensorExp(comp.kets[0], 34)

-- ``qset`` builds a ``ShardedRegister`` instead of a single-GPU ``DeviceState``: 2^n / P amplitudes per
GPU (``qbot_b200.sharded.ShardedKet``), and ``gate`` / ``swap`` / ``peek`` work on it through the same op
functions as on any other register:

    gate   queued on the sharded ket; fused sweeps per shard, NVLink exchange when a target is a rank bit
    swap   a relabelling of two qubits (no data moves)
    peek   outcome weights: local binning + one all-reduce; every rank gets the same ProbVal / result
    disc   what is left, Tr_rest psi psi^dagger of at most 10 kept qubits (made local, per-rank partial sums, one
           all-reduce), becomes an ordinary single-GPU density-matrix register on every rank
    meas   refused like on any ket-mode register above 13 qubits: the reference's collapse is a mixed
           product state (measurement.py:160-165) that only a 4^n density matrix can hold
    ProbVal-valued gates / conditions: refused for the same reason (they leave a mixed state)

Nothing here touches a device directly; the shards go through ``include/qbot_b200.h``.
"""
from __future__ import annotations

import os
import sys
from typing import Iterable, List, Optional, Sequence

import numpy as np

from .sharded import ShardedKet

SHARD_MIN_QUBITS = 32           # BASELINE north_star: "Large kets (>= 32 qubits) ... are partitioned across the GPUs"


class ShardingContext:
    """Who the ranks are and where the shards live.  One per process."""

    def __init__(self, comm, device: Optional[int] = None, min_qubits: Optional[int] = None, exchange: str = 'p2p',
                 shard_factory=None, jit: Optional[int] = None):
        self.comm = comm
        self.device = device
        self.min_qubits = int(min_qubits if min_qubits is not None else os.environ.get('QBOT_B200_SHARD_MIN_QUBITS', SHARD_MIN_QUBITS))
        self.exchange = exchange
        self.shard_factory = shard_factory
        self.jit = jit
        self._idle: dict = {}            # nq -> [ShardedKet] whose register was dropped (buffers + peer mappings stay open)

    def acquire(self, nq: int) -> ShardedKet:
        pool = self._idle.get(nq)
        if pool:
            return pool.pop()
        # a register of another size is gone for good: give its HBM back before allocating
        for other in list(self._idle):
            for sk in self._idle.pop(other):
                sk.close()
        sk = ShardedKet(nq, self.comm, shard_factory=self.shard_factory, device=self.device, exchange=self.exchange)
        if self.jit is not None and hasattr(sk.shard, 'set_jit'):
            sk.shard.set_jit(self.jit)
        return sk

    def release(self, sk: ShardedKet):
        # local bookkeeping only (no collective: this runs from __del__)
        self._idle.setdefault(sk.nq, []).append(sk)

    def close(self):
        for pool in self._idle.values():
            for sk in pool:
                sk.close()
        self._idle = {}


_ctx: Optional[ShardingContext] = None
_auto_failed = False


def enable(comm=None, device: Optional[int] = None, **kw) -> ShardingContext:
    """Shard large kets over the ranks of `comm` (default: torch.distributed's world, NCCL)."""
    global _ctx
    if comm is None:
        from .sharded import TorchComm
        comm = TorchComm()
    if device is None and kw.get('shard_factory') is None:
        device = int(os.environ.get('LOCAL_RANK', '0'))
    _ctx = ShardingContext(comm, device, **kw)
    return _ctx


def disable():
    global _ctx
    if _ctx is not None:
        _ctx.close()
    _ctx = None


def context() -> Optional[ShardingContext]:
    """The active context; under torchrun (torch.distributed initialised with more than one rank,
    or QBOT_B200_SHARD=1 with RANK / WORLD_SIZE set) it is created on first use."""
    global _ctx, _auto_failed
    if _ctx is not None or _auto_failed:
        return _ctx
    # one process, no launcher: nothing to shard over -- and no reason to import torch (seconds on a cold start;
    # the single-GPU path never needs it)
    if int(os.environ.get('WORLD_SIZE', '1') or 1) < 2 and 'torch.distributed' not in sys.modules:
        return None
    try:
        import torch.distributed as dist
        if not dist.is_available():
            return None
        if not dist.is_initialized():
            if os.environ.get('QBOT_B200_SHARD', '0') != '1' or int(os.environ.get('WORLD_SIZE', '1')) < 2:
                return None
            import torch
            local = int(os.environ.get('LOCAL_RANK', '0'))
            torch.cuda.set_device(local)
            dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
        if dist.get_world_size() < 2 or dist.get_backend() != 'nccl':
            return None
        return enable()
    except Exception:           # noqa: BLE001  (no torch / no NCCL: single-GPU registers only)
        _auto_failed = True
        return None


class ShardedRegister:
    """What the op functions (host/ops.py) see: the DeviceState methods they call, on a ShardedKet."""
    _qb_device_state = True
    _qb_sharded = True
    __array_priority__ = 1000
    kind = 0                     # KET
    nbranch = 1
    GATHER_MAX_QUBITS = 26       # user expressions may read `state` as an ndarray up to this size
    _version = 0                 # bumped by every update: a peek result's lazily computed rho_A checks it when it is read

    def __init__(self, sk: ShardedKet, ctx: ShardingContext):
        self._sk, self._ctx = sk, ctx
        self.nq = sk.nq
        self._shared = True

    @classmethod
    def product(cls, factors: Sequence[np.ndarray], ctx: ShardingContext) -> "ShardedRegister":
        sk = ctx.acquire(len(factors))
        sk.init_product([np.asarray(f, dtype=np.complex128).reshape(2) for f in factors])
        return cls(sk, ctx)

    def __del__(self):
        sk, ctx = getattr(self, '_sk', None), getattr(self, '_ctx', None)
        if sk is not None and ctx is not None:
            self._sk = None
            try:
                sk.queue = []
                ctx.release(sk)
            except Exception:       # noqa: BLE001
                pass

    # ---- ndarray contract (operators.py:14-17, helpers.py:24-37) ---------------------------------
    @property
    def shape(self):
        return (1 << self.nq,)

    ndim = 1

    @property
    def size(self):
        return 1 << self.nq

    @property
    def dtype(self):
        return np.dtype(np.complex128)

    def __array__(self, dtype=None, copy=None):
        if self.nq > self.GATHER_MAX_QUBITS:
            raise ValueError(f"the {self.nq}-qubit sharded register cannot be read as an array (limit {self.GATHER_MAX_QUBITS} qubits)")
        a = self._sk.gather()
        return a if dtype is None else a.astype(dtype)

    def __repr__(self):
        return f"ShardedRegister({self.nq} qubits over {self._sk.comm.world} ranks)"

    # ---- what gate / swap / peek call --------------------------------------------------------------
    def apply_gate(self, matrix, first_target: int = 0, controls: Iterable[int] = ()):
        self._sk.apply_gate(matrix, first_target, controls)
        self._version += 1
        return self

    def apply_swap(self, qubit_a: int, qubit_b: int):
        self._sk.swap_qubits(qubit_a, qubit_b)
        self._version += 1
        return self

    def probs(self, qubits: Sequence[int]) -> np.ndarray:
        return self._sk.probs(list(qubits))

    def probs_basis(self, qubits: Sequence[int], basis_kets) -> np.ndarray:
        """Outcome weights in a product of orthonormal measurement bases (measurement.py:88-101,
        147-155): rotate the groups by the basis kets, bin, rotate back (the register keeps its value up
        to rounding; a scratch copy of a sharded ket does not fit)."""
        w = np.ascontiguousarray(np.stack([np.asarray(k, dtype=np.complex128).reshape(-1) for k in basis_kets]))
        b = int(w.shape[1]).bit_length() - 1
        if w.shape != (1 << b, 1 << b):
            raise ValueError("a measurement basis needs 2^b kets of 2^b amplitudes")
        if not np.allclose(w @ w.conj().T, np.eye(1 << b), atol=1e-12):
            raise ValueError("measurement on a sharded register needs an orthonormal basis")
        qs = list(qubits)
        groups = [qs[i:i + b] for i in range(0, len(qs), b)]
        for gq in groups:
            self._sk.apply_gate_qubits(w, gq)
        out = self._sk.probs(qs)
        winv = np.ascontiguousarray(w.conj().T)
        for gq in groups:
            self._sk.apply_gate_qubits(winv, gq)
        return out

    def ptrace_keep(self, keep_qubits: Sequence[int]) -> np.ndarray:
        """rho_A of the listed qubits (what a peek result's unMeasuredDensity is), as a host array."""
        return self._sk.reduced_density(list(keep_qubits))

    def clone(self):
        raise ValueError("a sharded register cannot be copied (a program that names `state` needs a single-GPU register)")

    def as_density(self):
        raise ValueError(f"the {self.nq}-qubit sharded ket cannot become a density matrix")

    def flush(self):
        self._sk.flush()

    def sync(self):
        self._sk.sync()

    def norm2(self) -> float:
        return self._sk.norm2()

    def amplitudes(self, indices: Sequence[int]) -> np.ndarray:
        return self._sk.amplitudes(indices)

    def stats(self) -> dict:
        st = getattr(self._sk.shard, 'state', None)
        out = dict(st.stats()) if st is not None else {}
        out['exchanges'] = self._sk.shard.exchanges
        out['exchanged_bytes'] = self._sk.shard.exchanged_bytes
        return out
