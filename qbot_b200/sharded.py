"""Kets sharded over the GPUs of one box (BASELINE config 5).

One process per GPU.  A ket of n qubits lives as P = 2^g contiguous slices: the top g index
bits select the rank (physical positions nl .. n-1, nl = n - g), the low nl bits address the
amplitude inside the rank's shard.  Which *logical* index bit sits at which physical position
is a permutation owned by ``QubitMap`` and is the same on every rank.

  - gates whose non-diagonal targets are all local run on the shard through the C ABI (fused
    tile engine); a control on a rank bit is a rank predicate, a diagonal factor on a rank bit
    is a per-rank scalar -- no communication;
  - when a gate needs a non-diagonal target that currently selects the rank, ``ShardedKet``
    exchanges qubits: it picks the logical bits whose next non-diagonal use is farthest away
    to become the new rank bits, packs the shard so that the outgoing bits are the top local
    bits (one index-bit permutation pass) and swaps chunks with the 2^k - 1 peers.  With
    ``exchange='p2p'`` the pack kernel stores straight into the peers' buffers over NVLink
    (CUDA IPC mappings; pack and exchange are ONE kernel); with ``exchange='nccl'`` the pack
    goes to the local second buffer and a grouped all-to-all (torch.distributed / NCCL) moves
    the chunks.  No swap-back ever happens: the map just changes.

The reference has no counterpart (it is single-process numpy); the operation replaced is the
``gate`` op's state path (qbot/operators.py:255-329 -> qgates.py:161-182, 228-279) at sizes the
reference cannot represent.  Everything in this file is host logic; device work goes through
``include/qbot_b200.h``.
"""
from __future__ import annotations

import ctypes as C
import os
import time
from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

INF = 1 << 60


def _trace(msg: str):
    if os.environ.get('QBOT_B200_TRACE'):
        import sys
        print(f"[qbot_b200 rank {os.environ.get('RANK', '?')}] {msg}", file=sys.stderr, flush=True)
LOW_KEEP = 5          # local positions 0..4 stay in place during a pack (512-byte runs)


# ---------------------------------------------------------------------------------------------
# gate records on logical index bits
# ---------------------------------------------------------------------------------------------
@dataclass
class LGate:
    """A gate on logical index bits: ``tb[0]`` carries the matrix's most significant bit."""
    m: np.ndarray
    tb: Tuple[int, ...]
    controls: Tuple[int, ...] = ()
    wmask: int = 0        # logical bits the gate acts on non-diagonally
    rmask: int = 0        # logical bits it only reads (controls, block-diagonal target axes)


def _axis_is_diagonal(m: np.ndarray, k: int, axis: int) -> bool:
    """True if m never couples basis states that differ in target `axis` (0 = most significant)."""
    d = 1 << k
    bit = 1 << (k - 1 - axis)
    idx = np.arange(d)
    off = (idx[:, None] & bit) != (idx[None, :] & bit)
    return not np.any(m[off] != 0)


def make_lgate(matrix, target_bits: Sequence[int], control_bits: Iterable[int] = ()) -> LGate:
    m = np.ascontiguousarray(np.asarray(matrix), dtype=np.complex128)
    k = len(target_bits)
    if m.shape != (1 << k, 1 << k):
        raise ValueError("matrix does not match the number of target bits")
    controls = tuple(int(c) for c in control_bits)
    if set(controls) & set(target_bits):
        raise ValueError("control overlaps target")
    w = r = 0
    if k == 1:            # (the common case, without numpy index arithmetic: 4 us instead of 20 per gate of a queued circuit)
        if m[0, 1] == 0 and m[1, 0] == 0:
            r = 1 << target_bits[0]
        else:
            w = 1 << target_bits[0]
    else:
        for ax, b in enumerate(target_bits):
            if _axis_is_diagonal(m, k, ax):
                r |= 1 << b
            else:
                w |= 1 << b
    for c in controls:
        r |= 1 << c
    return LGate(m, tuple(int(b) for b in target_bits), controls, w, r)


def select_pass(rem: Sequence[LGate], allowed: int) -> List[int]:
    """Indices of the gates of `rem` (circuit order) that can run now when only the bits of
    `allowed` may be written: a gate may overtake an earlier gate left behind only if neither
    writes a bit the other touches (same rule as the fusion planner, qb_plan.cpp)."""
    picked = []
    blocked_w = blocked_r = 0
    for i, g in enumerate(rem):
        ok = (g.wmask & ~allowed) == 0 and ((g.wmask | g.rmask) & blocked_w) == 0 and (g.wmask & blocked_r) == 0
        if ok:
            picked.append(i)
        else:
            blocked_w |= g.wmask
            blocked_r |= g.rmask
    return picked


# ---------------------------------------------------------------------------------------------
# qubit map + exchange planning (pure host logic, identical on every rank)
# ---------------------------------------------------------------------------------------------
@dataclass
class Exchange:
    """One global-qubit swap.  `src_bit_of_dst_bit` is the local pack permutation; `rank_bits`
    the rank-bit numbers (ascending) whose qubits come in; chunk-index bit i pairs with
    rank_bits[i]."""
    src_bit_of_dst_bit: List[int]
    rank_bits: List[int]
    hbits: Tuple[int, ...] = ()     # logical bits parked at the top local positions: the sub-blocks of a pipelined exchange
    tail: Tuple[int, ...] = ()      # the parked bits were on top before the exchange too: these gates of the pass before it (indices) can run
                                    # last and leave the parked bits alone -- their sweeps run sub-block by sub-block, each followed by its piece

    @property
    def send_side(self) -> bool:
        return len(self.tail) > 0

    @property
    def k(self) -> int:
        return len(self.rank_bits)

    @property
    def split(self) -> int:
        """v: the shard is exchanged in 2^v pieces (packed layout [v piece bits][k chunk bits][rest]); 0 = in one go"""
        return len(self.hbits)


class QubitMap:
    def __init__(self, n: int, g: int):
        if g < 0 or g > n:
            raise ValueError("bad number of rank bits")
        self.n, self.g, self.nl = n, g, n - g
        self.pos = list(range(n))        # logical bit -> physical position
        self.at = list(range(n))         # physical position -> logical bit

    def is_local(self, b: int) -> bool:
        return self.pos[b] < self.nl

    def local_mask(self) -> int:
        m = 0
        for b in range(self.n):
            if self.pos[b] < self.nl:
                m |= 1 << b
        return m

    def choose_initial(self, rem: Sequence[LGate]):
        """A register that is still a PRODUCT state can start with any qubits on the rank bits (a relabelling, no data
        moves): take the logical bits whose first non-diagonal use among the queued gates `rem` is farthest away (never
        written first; ties keep the current rank bits), the same rule plan_exchange applies later.  With the identity
        map the first exchange of rc(34, 10) on 8 ranks comes after 141 gates and a second one is needed for the
        last gate; chosen this way it comes after 215 and is the only one."""
        n, nl, g = self.n, self.nl, self.g
        nxt = [INF] * n
        for i, gt in enumerate(rem):
            w = gt.wmask
            while w:
                b = (w & -w).bit_length() - 1
                if nxt[b] == INF:
                    nxt[b] = i
                w &= w - 1
        order = sorted(range(n), key=lambda b: (-nxt[b], 0 if self.pos[b] >= nl else 1, -self.pos[b]))
        new_global = order[:g]
        cur_global = [self.at[nl + r] for r in range(g)]
        incoming = [b for b in new_global if b not in cur_global]
        outgoing = [b for b in cur_global if b not in new_global]
        for bi, bo in zip(incoming, outgoing):
            pi, po = self.pos[bi], self.pos[bo]
            self.at[pi], self.at[po] = bo, bi
            self.pos[bi], self.pos[bo] = po, pi

    def plan_exchange(self, rem: Sequence[LGate], split: int = 0, min_first_phase: int = 40,
                      prev: Optional[Sequence[LGate]] = None, phase_cap: int = 96) -> Exchange:
        """Choose the new rank bits (farthest next non-diagonal use) and update the map.  `prev`: the gates of the
        pass that runs right before this exchange, not applied yet (for the sending-side overlap, see below; the
        caller localises them with the map as it is BEFORE this call).

        split = v > 0 asks for a PIPELINED exchange: v local bits are parked at the top v local positions, so that
        the shard is 2^v contiguous sub-blocks that are exchanged one after the other while the sweeps next to the
        exchange run sub-block by sub-block -- the last sweeps of the pass before it (each piece leaves as soon as
        its sub-block is done) and the first sweeps of the pass after it (each sub-block starts as soon as its piece
        has arrived).  A sweep can run that way when the parked bits are not tile bits of it, i.e. when none of its
        gates writes them, so the parked bits are either (A) the ones already on top, if few of the gates around the
        exchange write them (both sides overlap), or (B) the bits written last among the gates that follow (only the
        receiving side overlaps: the sub-blocks do not exist before this exchange).  The count of gates that could
        run in the overlapped phases (capped at phase_cap each) decides; below min_first_phase the exchange is
        done in one go (hbits = ())."""
        n, nl, g = self.n, self.nl, self.g
        nxt = [INF] * n
        for i, gt in enumerate(rem):
            w = gt.wmask
            while w:
                b = (w & -w).bit_length() - 1
                if nxt[b] == INF:
                    nxt[b] = i
                w &= w - 1
        # a local bit in the low LOW_KEEP positions is a last resort (keeps the pack coalesced)
        # (ties: stay global > ordinary local positions, highest first > the top `split` local positions, where the
        # bits parked by a pipelined exchange sit: they are worth more as the sub-block bits of the next one)
        order = sorted(range(n), key=lambda b: (0 if not (self.pos[b] < min(LOW_KEEP, max(nl - g, 0))) else 1,
                                                -nxt[b], 0 if self.pos[b] >= nl else 2 if self.pos[b] >= nl - split else 1,
                                                -self.pos[b]))
        head_w = rem[0].wmask if rem else 0
        new_global = []
        for b in order:
            if (head_w >> b) & 1:
                continue
            new_global.append(b)
            if len(new_global) == g:
                break
        if len(new_global) < g:
            raise RuntimeError("cannot make the head gate local: it writes more bits than a shard holds")
        cur_global = [self.at[nl + r] for r in range(g)]
        victims = sorted((b for b in new_global if b not in cur_global), key=lambda b: self.pos[b])
        rank_bits = sorted(r for r in range(g) if self.at[nl + r] not in new_global)
        k = len(victims)
        assert k == len(rank_bits)
        hbits: List[int] = []
        tail: List[int] = []
        if split > 0 and k and nl - k - split >= max(LOW_KEEP, 8):
            def first_phase(hb):
                allowed = 0
                for b in range(n):
                    if b not in new_global and b not in hb:
                        allowed |= 1 << b
                return min(len(select_pass(rem, allowed)), phase_cap)

            # (B) park the bits that are written last among the gates that follow: overlap on the receiving side
            cand_b: List[int] = []
            for b in order:                     # farthest next write first
                if b in new_global or self.pos[b] >= nl or self.pos[b] < LOW_KEEP or (head_w >> b) & 1:
                    continue
                cand_b.append(b)
                if len(cand_b) == split:
                    break
            score_b = first_phase(cand_b) if len(cand_b) == split else -1
            # (A) keep the bits that sit at the top local positions now: the sub-blocks exist on the SENDING side too,
            # so the tail of the pass before the exchange (the gates of `prev` that can run last and leave those bits
            # alone) runs sub-block by sub-block, each followed at once by its piece of the exchange
            cand_a = [self.at[nl - split + i] for i in range(split)]
            score_a, tail_a = -1, []
            if prev and not any(b in new_global or (head_w >> b) & 1 for b in cand_a):
                hm = 0
                for b in cand_a:
                    hm |= 1 << b
                lm = self.local_mask()
                rev = select_pass(list(reversed(prev)), lm & ~hm)          # gates that commute to the END of the pass
                # (only as many gates as it takes to cover the exchange: sub-block sweeps share the SMs with the
                # scatter kernels; any suffix of the deferrable set is itself deferrable)
                tail_a = sorted(len(prev) - 1 - i for i in rev)[-phase_cap:]
                score_a = len(tail_a) + first_phase(cand_a)
            if score_a >= score_b and score_a >= min_first_phase:
                hbits, tail = cand_a, tail_a
            elif score_b >= min_first_phase:
                hbits = cand_b
        v = len(hbits)
        perm = list(range(nl))
        if k:
            # packed layout, top down: [v parked bits][k outgoing bits][the rest]
            top = list(range(nl - v - k, nl))
            want = [self.pos[b] for b in victims] + [self.pos[b] for b in hbits]
            displaced = [p for p in top if p not in want]          # others sitting in the top slots
            vacated = [p for p in want if p not in top]            # wanted bits' slots outside the top
            for i, p in enumerate(want):
                perm[nl - v - k + i] = p
            for d, s in zip(vacated, displaced):
                perm[d] = s
            # new map: pack first ...
            new_at = list(self.at)
            for d in range(nl):
                new_at[d] = self.at[perm[d]]
            # ... then the exchange swaps chunk-index position i with rank bit rank_bits[i]
            for i, r in enumerate(rank_bits):
                new_at[nl - v - k + i], new_at[nl + r] = new_at[nl + r], new_at[nl - v - k + i]
            self.at = new_at
            for p, b in enumerate(self.at):
                self.pos[b] = p
        return Exchange(perm, rank_bits, tuple(hbits), tuple(tail))

    def localise(self, g: LGate, rank: int, nl: Optional[int] = None):
        """The gate as this rank sees it: (matrix, local target positions, local control mask),
        or None when a rank-bit control is 0 here.  Block-diagonal target axes that sit on rank
        bits are sliced by the rank's bit value.  With nl < self.nl the top local positions count as
        rank bits too (sub-blocks of a pipelined exchange: rank = (rank << v) | sub-block number)."""
        nl = self.nl if nl is None else nl
        cmask = 0
        for c in g.controls:
            p = self.pos[c]
            if p >= nl:
                if not (rank >> (p - nl)) & 1:
                    return None
            else:
                cmask |= 1 << p
        k = len(g.tb)
        m = g.m
        glob = [(ax, self.pos[b]) for ax, b in enumerate(g.tb) if self.pos[b] >= nl]
        if glob:
            t = m.reshape([2] * (2 * k))
            keep_axes = [ax for ax in range(k) if self.pos[g.tb[ax]] < nl]
            sl = [slice(None)] * (2 * k)
            for ax, p in glob:
                if (g.wmask >> g.tb[ax]) & 1:
                    raise RuntimeError("localise: non-diagonal target on a rank bit (exchange missing)")
                v = (rank >> (p - nl)) & 1
                sl[ax] = v
                sl[k + ax] = v
            t = t[tuple(sl)]
            kk = len(keep_axes)
            if kk == 0:
                s = complex(t)
                # a per-rank scalar: diag(s, s) on a local bit outside the control mask
                free = next((p for p in range(nl) if not (cmask >> p) & 1), None)
                if free is None:
                    # every local bit is a control: diag(1, s) on one of them, the others stay controls
                    p = (cmask & -cmask).bit_length() - 1
                    return np.array([[1, 0], [0, s]], dtype=np.complex128), [p], cmask & ~(1 << p)
                return np.array([[s, 0], [0, s]], dtype=np.complex128), [free], cmask
            m = np.ascontiguousarray(t.reshape(1 << kk, 1 << kk))
            tbits = [self.pos[g.tb[ax]] for ax in keep_axes]
            return m, tbits, cmask
        return m, [self.pos[b] for b in g.tb], cmask


# ---------------------------------------------------------------------------------------------
# communicators
# ---------------------------------------------------------------------------------------------
class TorchComm:
    """torch.distributed plumbing (NCCL on the GPUs, gloo in the CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def barrier(self):
        self.dist.barrier(group=self.group)

    def allreduce_sum(self, arr: np.ndarray, device=None) -> np.ndarray:
        import torch
        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64))
        if device is not None:
            t = t.to(device)
        self.dist.all_reduce(t, group=self.group)
        return t.cpu().numpy()

    def allgather_bytes(self, payload: bytes) -> List[bytes]:
        out = [None] * self.world
        self.dist.all_gather_object(out, payload, group=self.group)
        return out

    def all_to_all(self, out_t, in_t, out_splits, in_splits):
        self.dist.all_to_all_single(out_t, in_t, out_splits, in_splits, group=self.group)


class _CudaArray:
    """Minimal __cuda_array_interface__ carrier so that torch can wrap a raw device pointer."""

    def __init__(self, ptr: int, nfloat64: int):
        self.__cuda_array_interface__ = {'shape': (nfloat64,), 'typestr': '<f8', 'data': (ptr, False), 'version': 2}


# ---------------------------------------------------------------------------------------------
# the shard on a GPU
# ---------------------------------------------------------------------------------------------
class CudaShard:
    """2^nl amplitudes in HBM (two buffers: the live shard and the exchange target)."""

    def __init__(self, nl: int, comm, device: int, exchange: str = 'p2p'):
        from . import _lib
        from .state import DeviceState, KET
        self._lib = _lib
        self.nl, self.comm, self.device = nl, comm, device
        self.bytes = 16 << nl
        self.buf = [C.c_void_p(), C.c_void_p()]
        for b in self.buf:
            _lib.call('qb_buffer_alloc', device, self.bytes, C.byref(b))
        self.cur = 0
        h = C.c_void_p()
        _lib.call('qb_create_external', C.byref(h), KET, nl, 1, device, self.buf[0], None)
        self.state = DeviceState(h, KET, nl, 1)
        self.exchange_mode = exchange
        self.peer = None                      # peer[r][i] = mapping of rank r's buffer i
        self.peer_flags = None                # peer_flags[r] = mapping of rank r's flag block (one 8-byte counter per source rank)
        self.flags = C.c_void_p()
        self.exchanged_bytes = 0
        self.exchanges = 0
        self.split_exchanges = 0
        self.exchange_seconds = 0.0           # pack + transfer: between the two barriers, or (pipelined) first piece start -> last piece done on the exchange stream
        self._pending_times = []              # (start event, end event) of pipelined exchanges not yet read
        self._incoming = None                 # pieces of a pipelined exchange the next sweeps have to wait for
        self._copy_streams = None
        self._stage_busy = None               # event: the copy engines have read the staging buffer (pieces of the last exchange)
        # who carries the pieces of a pipelined exchange: 'ce' = the copy engines (a local pack pass, then plain copies
        # into the peers' live buffers; the sweeps keep every SM), 'sm' = peer stores issued by a scatter kernel
        self.piece_mode = os.environ.get('QBOT_B200_EXCHANGE_PIECES', 'sm')      # (measured at 8 GPUs: 'sm' +3.5 %, 'ce' -30 % against the plain exchange)
        self._sms = None
        self.overlapped_steps = 0             # sweeps run sub-block by sub-block beside an exchange
        # sweeps on either side of a pipelined exchange that run sub-block by sub-block (about what covers the exchange)
        self.overlap_steps = int(os.environ.get('QBOT_B200_EXCHANGE_OVERLAP_STEPS', '4'))
        self._jit = None
        self._xseq = 0                        # pieces signalled so far (same on every rank)
        self._streams = None
        # SM slots a pipelined exchange's scatter kernels may take (256-thread CTAs); the sweeps that run beside them
        # size their persistent grids for the rest
        self.exchange_ctas = int(os.environ.get('QBOT_B200_EXCHANGE_CTAS', '48'))
        if exchange == 'p2p' and comm.world > 1:
            _lib.call('qb_buffer_alloc', device, 4096, C.byref(self.flags))
            import torch
            torch.as_tensor(_CudaArray(self.flags.value, 512), device=f'cuda:{device}').zero_()
            torch.cuda.synchronize(device)
            self._open_peers()
        elif exchange not in ('p2p', 'nccl'):
            raise ValueError("exchange must be 'p2p' or 'nccl'")

    @property
    def supports_split(self) -> bool:
        return self.peer is not None

    def timer_start(self):
        self.state.timer_start()

    def timer_stop(self) -> float:
        return self.state.timer_stop()       # the compute stream: sub-block sweeps and the waits for the exchange pieces are on it

    def reset_stats(self):
        self.state.reset_stats()

    def stats(self) -> dict:
        """Counters of the shard handle (a sweep run sub-block by sub-block counts as one pass, with its last part)."""
        return {k: float(x) for k, x in self.state.stats().items()}

    def set_jit(self, mode: int):
        self._jit = mode
        self.state.set_jit(mode)

    def _open_peers(self):
        lib = self._lib
        handles = b''
        for b in self.buf + [self.flags]:
            hb = C.create_string_buffer(64)
            lib.call('qb_ipc_export', self.device, b, hb)
            handles += hb.raw
        allh = self.comm.allgather_bytes(handles)
        self.peer, self.peer_flags = [], []
        for r, hs in enumerate(allh):
            if r == self.comm.rank:
                self.peer.append([self.buf[0].value, self.buf[1].value])
                self.peer_flags.append(self.flags.value)
                continue
            ptrs = []
            for i in range(3):
                p = C.c_void_p()
                lib.call('qb_ipc_open', self.device, C.create_string_buffer(hs[64 * i:64 * i + 64], 64), C.byref(p))
                ptrs.append(p.value)
            self.peer.append(ptrs[:2])
            self.peer_flags.append(ptrs[2])

    def _stream_pair(self):
        """(compute stream of the library as a torch stream, high-priority exchange stream)"""
        if self._streams is None:
            import torch
            p = C.c_void_p()
            self._lib.call('qb_compute_stream', self.device, C.byref(p))
            cs = torch.cuda.ExternalStream(p.value, device=f'cuda:{self.device}')
            xs = torch.cuda.Stream(device=f'cuda:{self.device}', priority=-1)
            tok = torch.zeros(1, dtype=torch.float32, device=f'cuda:{self.device}')
            self._streams = (cs, xs, tok)
        return self._streams

    def _plan_queue(self, v: int):
        n, hd, tl = C.c_int(0), C.c_int(0), C.c_int(0)
        self._lib.call('qb_plan_queue', self.state._h, v, C.byref(n), C.byref(hd), C.byref(tl))
        return n.value, hd.value, tl.value

    def _run_queued(self, send_v: int = 0, on_part_done=None, sm_limit: Optional[int] = None):
        """Run the queued pass.  Its first sweeps wait, sub-block by sub-block, for the pieces of a pipelined
        exchange that is still arriving (self._incoming); its last sweeps run sub-block by sub-block when the
        pieces of the NEXT exchange are to leave behind them (send_v > 0; on_part_done(j) is called after sub-block
        j's last sweep has been queued).  Everything else runs over the whole shard."""
        lib, inc = self._lib, self._incoming
        v = inc['v'] if inc is not None else send_v
        if inc is not None and send_v and send_v != inc['v']:
            send_v = 0                           # (different sub-block sizes on the two sides: keep the receiving one)
        n, head, tail = self._plan_queue(v) if v else self._plan_queue(0)
        cs = self._stream_pair()[0] if (inc is not None or send_v) else None
        h = min(head, self.overlap_steps) if inc is not None else 0
        t = min(tail, self.overlap_steps, n - h) if send_v else 0
        Q = 1 << v
        sm = self._sm_limit() if sm_limit is None else sm_limit         # sweeps beside SM-issued peer stores leave SMs to them
        sm_in = inc.get('sm_limit', self._sm_limit()) if inc is not None else 0
        if inc is not None:
            for j in range(Q):
                ev, seq = inc['arrived'][j]
                cs.wait_event(ev)                 # my own piece j has left ...
                lib.call('qb_wait_flags', self.device, None, inc['wait'], inc['nwait'], seq)       # ... and every source's piece j is in
                if h:
                    lib.call('qb_run_steps', self.state._h, 0, h, j, Q, sm_in)
            self._incoming = None
            self.overlapped_steps += h
        if n - h - t > 0:
            lib.call('qb_run_steps', self.state._h, h, n - t, 0, 1, 0)
        if send_v:
            for j in range(Q):
                if t:
                    lib.call('qb_run_steps', self.state._h, n - t, n, j, Q, sm)
                on_part_done(j)
            self.overlapped_steps += t
        if n:
            lib.call('qb_finish_queue', self.state._h)

    def _sm_limit(self) -> int:
        if self._sms is None:
            sms = C.c_int(0)
            self._lib.call('qb_device_info', self.device, None, 0, C.byref(sms), None, None, None)
            self._sms = sms.value
        return max(self._sms - (self.exchange_ctas + 1) // 2, 8)

    def do_exchange_split(self, ex: "Exchange"):
        """The exchange in 2^v pieces on a stream of its own.  Sending side (ex.send_side): the last sweeps of the
        queued pass run sub-block by sub-block and piece j leaves right behind sub-block j.  Receiving side: the
        first sweeps of the NEXT pass run on sub-block j as soon as piece j has arrived from every source (see
        _run_queued).  No host synchronisation: pieces are ordered by events (local) and by flag counters in the
        receivers' memory (between GPUs)."""
        import torch
        lib, comm = self._lib, self.comm
        k, v, nl = ex.k, ex.split, self.nl
        Q, sub_bits = 1 << v, nl - v
        rank = comm.rank
        my_s = 0
        for i, r in enumerate(ex.rank_bits):
            my_s |= ((rank >> r) & 1) << i
        sub_bytes = self.bytes >> v
        chunk_bytes = sub_bytes >> k
        other = 1 - self.cur

        def peer_rank(c):
            pr = rank
            for i, r in enumerate(ex.rank_bits):
                pr = (pr & ~(1 << r)) | (((c >> i) & 1) << r)
            return pr

        group = [peer_rank(c) for c in range(1 << k) if peer_rank(c) != rank]
        cs, xs, tok = self._stream_pair()
        perm = ex.src_bit_of_dst_bit
        perm_sub = lib.int_array(perm[:sub_bits])
        fixed_mask = 0
        for i in range(v):
            fixed_mask |= 1 << perm[sub_bits + i]
        dst = (C.c_void_p * (1 << k))()
        sig = (C.c_void_p * max(len(group), 1))(*[self.peer_flags[p] + 8 * rank for p in group])
        wait = (C.c_void_p * max(len(group), 1))(*[self.flags.value + 8 * p for p in group])
        if self.piece_mode == 'ce':
            return self._exchange_split_ce(ex, peer_rank, group, my_s)
        arrived = []
        started = []
        with torch.cuda.stream(xs):
            # Every rank is past the scatters of the previous exchange (they read the buffers that are about to be
            # overwritten remotely; those buffers have not been touched since): one tiny all-reduce on the exchange
            # streams, queued before the tail sweeps so that it is long done when the first piece is ready to leave
            comm.dist.all_reduce(tok, group=comm.group)

        def piece(j):
            done = torch.cuda.Event()
            done.record(cs)                        # sub-block j of the pass before the exchange is finished
            xs.wait_event(done)
            if not started:
                t0 = torch.cuda.Event(enable_timing=True)
                t0.record(xs)
                started.append(t0)
            fixed_val = 0
            for i in range(v):
                fixed_val |= ((j >> i) & 1) << perm[sub_bits + i]
            for c in range(1 << k):
                dst[c] = self.peer[peer_rank(c)][other] + j * sub_bytes + my_s * chunk_bytes
            lib.call('qb_permute_scatter_sub', self.state._h, perm_sub, sub_bits, fixed_mask, fixed_val, k, dst, my_s,
                     C.c_void_p(xs.cuda_stream), self.exchange_ctas)
            self._xseq += 1
            lib.call('qb_signal_flags', self.device, C.c_void_p(xs.cuda_stream), sig, len(group), self._xseq)
            ev = torch.cuda.Event()
            ev.record(xs)
            arrived.append((ev, self._xseq))

        self._run_queued(send_v=v if ex.send_side else 0, on_part_done=piece if ex.send_side else None)
        if not ex.send_side:
            for j in range(Q):
                piece(j)
        t1 = torch.cuda.Event(enable_timing=True)
        t1.record(xs)
        self._pending_times.append((started[0], t1))
        lib.call('qb_rebind', self.state._h, self.buf[other])
        self.cur = other
        self._incoming = dict(v=v, arrived=arrived, wait=wait, nwait=len(group))
        self.exchanged_bytes += (self.bytes >> k) * ((1 << k) - 1)
        self.exchanges += 1
        self.split_exchanges += 1

    def _exchange_split_ce(self, ex, peer_rank, group, my_s):
        """Pieces carried by the COPY ENGINES: sub-block j is packed into the second buffer (one local pass over the
        sub-block, on the compute stream, right behind the sub-block's last sweeps), and its chunks then travel as
        plain device-to-device copies into the peers' LIVE buffers -- no SM takes part, so the sweeps of the other
        sub-blocks keep the whole GPU.  A receiver's sub-block may be overwritten once the receiver has packed it
        (`packed` counters), and is complete once every source has delivered (`arrived` counters).  The live buffer
        stays the live buffer."""
        import torch
        lib, rank = self._lib, self.comm.rank
        k, v, nl = ex.k, ex.split, self.nl
        Q, sub_bits = 1 << v, nl - v
        sub_bytes = self.bytes >> v
        chunk_bytes = sub_bytes >> k
        live, stage = self.buf[self.cur].value, self.buf[1 - self.cur].value
        cs, xs, _ = self._stream_pair()
        perm = ex.src_bit_of_dst_bit
        perm_sub = lib.int_array(perm[:sub_bits])
        fixed_mask = 0
        for i in range(v):
            fixed_mask |= 1 << perm[sub_bits + i]
        ng = len(group)
        sig_packed = (C.c_void_p * max(ng, 1))(*[self.peer_flags[p] + 16 * rank for p in group])
        sig_arrived = (C.c_void_p * max(ng, 1))(*[self.peer_flags[p] + 16 * rank + 8 for p in group])
        wait_packed = (C.c_void_p * max(ng, 1))(*[self.flags.value + 16 * p for p in group])
        wait_arrived = (C.c_void_p * max(ng, 1))(*[self.flags.value + 16 * p + 8 for p in group])
        xstream = C.c_void_p(xs.cuda_stream)
        if self._copy_streams is None:
            self._copy_streams = [torch.cuda.Stream(device=f'cuda:{self.device}', priority=-1) for _ in range(4)]
        copy_streams = self._copy_streams
        dst = (C.c_void_p * (1 << k))()
        arrived, started = [], []
        if self._stage_busy is not None:
            cs.wait_event(self._stage_busy)        # the previous exchange's copies have left the staging buffer
            self._stage_busy = None

        def piece(j):
            fixed_val = 0
            for i in range(v):
                fixed_val |= ((j >> i) & 1) << perm[sub_bits + i]
            for c in range(1 << k):
                dst[c] = stage + j * sub_bytes + c * chunk_bytes
            lib.call('qb_permute_scatter_sub', self.state._h, perm_sub, sub_bits, fixed_mask, fixed_val, k, dst, 0, None, 0)
            packed = torch.cuda.Event()
            packed.record(cs)
            xs.wait_event(packed)
            if not started:
                t0 = torch.cuda.Event(enable_timing=True)
                t0.record(xs)
                started.append(t0)
            self._xseq += 1
            seq = self._xseq
            lib.call('qb_signal_flags', self.device, xstream, sig_packed, ng, seq)       # my sub-block j may be overwritten
            lib.call('qb_wait_flags', self.device, xstream, wait_packed, ng, seq)        # ... and so may the peers'
            # the chunks of a piece go to different peers: several copy streams, so that several copy engines carry them
            go = torch.cuda.Event()
            go.record(xs)
            for c in range(1 << k):
                pr = peer_rank(c)
                st_c = copy_streams[c % len(copy_streams)]
                st_c.wait_event(go)
                lib.call('qb_copy_async', self.device, C.c_void_p(self.peer[pr][self.cur] + j * sub_bytes + my_s * chunk_bytes),
                         C.c_void_p(stage + j * sub_bytes + c * chunk_bytes), chunk_bytes, C.c_void_p(st_c.cuda_stream))
            for st_c in copy_streams[:min(len(copy_streams), 1 << k)]:
                fin = torch.cuda.Event()
                fin.record(st_c)
                xs.wait_event(fin)
            lib.call('qb_signal_flags', self.device, xstream, sig_arrived, ng, seq)
            ev = torch.cuda.Event()
            ev.record(xs)
            arrived.append((ev, seq))

        self._run_queued(send_v=v if ex.send_side else 0, on_part_done=piece if ex.send_side else None, sm_limit=0)
        if not ex.send_side:
            for j in range(Q):
                piece(j)
        t1 = torch.cuda.Event(enable_timing=True)
        t1.record(xs)
        self._pending_times.append((started[0], t1))
        self._stage_busy = t1
        self._incoming = dict(v=v, arrived=arrived, wait=wait_arrived, nwait=ng, sm_limit=0)
        self.exchanged_bytes += (self.bytes >> k) * ((1 << k) - 1)
        self.exchanges += 1
        self.split_exchanges += 1

    def _collect_times(self):
        for t0, t1 in self._pending_times:
            t1.synchronize()
            self.exchange_seconds += t0.elapsed_time(t1) / 1e3
        self._pending_times = []

    def init_basis(self, has_one: bool, local_index: int = 0):
        import torch
        self.flush()
        if has_one:
            self._lib.call('qb_init_basis', self.state._h, local_index)
        else:
            t = torch.as_tensor(_CudaArray(self.buf[self.cur].value, 2 << self.nl), device=f'cuda:{self.device}')
            t.zero_()
            torch.cuda.synchronize(self.device)
        self.state.sync()

    def init_product(self, local_factors: Sequence[np.ndarray], coeff: complex):
        """coeff * kron(local_factors) (first factor = highest local position) built on the device."""
        self.flush()
        v = np.ascontiguousarray(np.stack([np.asarray(f, dtype=np.complex128).reshape(2) for f in local_factors]))
        v[0] = v[0] * coeff
        self._lib.call('qb_init_product', self.state._h, v.ctypes.data_as(C.c_void_p), 0)
        self.state._dirty()
        self.state.sync()

    def rdm_local(self, positions: Sequence[int]) -> np.ndarray:
        """sum over the other local bits of psi psi^dagger on the listed local positions (first = MSB)."""
        nl = self.nl
        self.flush()
        return np.asarray(self.state.ptrace_keep([nl - 1 - p for p in positions]))

    def apply(self, m: np.ndarray, target_positions: Sequence[int], cmask: int):
        self.state.apply_gate_bits(m, list(target_positions), cmask)

    def flush(self):
        if self._incoming is not None:
            self._run_queued()
        else:
            self.state.flush()

    def sync(self):
        self.flush()
        self.state.sync()
        if self._pending_times:
            self._collect_times()
            n = C.c_uint64(0)
            self._lib.call('qb_flag_timeouts', self.device, C.byref(n))
            if n.value:
                raise RuntimeError(f"pipelined exchange: {n.value} wait(s) for a peer's piece timed out")

    def do_exchange(self, ex: Exchange):
        import torch
        lib, comm = self._lib, self.comm
        k, nl = ex.k, self.nl
        if k == 0:
            return
        rank = comm.rank
        my_s = 0
        for i, r in enumerate(ex.rank_bits):
            my_s |= ((rank >> r) & 1) << i
        chunk_bytes = self.bytes >> k
        other = 1 - self.cur

        def peer_rank(c):
            pr = rank
            for i, r in enumerate(ex.rank_bits):
                pr = (pr & ~(1 << r)) | (((c >> i) & 1) << r)
            return pr

        perm = lib.int_array(ex.src_bit_of_dst_bit)
        dst = (C.c_void_p * (1 << k))()
        if self.exchange_mode == 'p2p':
            # everyone must be done with the buffer we are about to overwrite remotely
            self.state.sync()
            comm.barrier()
            t0 = time.perf_counter()
            for c in range(1 << k):
                dst[c] = self.peer[peer_rank(c)][other] + my_s * chunk_bytes
            # chunk order rotated by this rank's own chunk number: a pairwise-exchange schedule
            lib.call('qb_permute_scatter', self.state._h, perm, k, dst, my_s)
            self.state.sync()
            self.exchange_seconds += time.perf_counter() - t0
            comm.barrier()
            lib.call('qb_rebind', self.state._h, self.buf[other])
            self.cur = other
        else:
            self.state.sync()
            t0 = time.perf_counter()
            for c in range(1 << k):
                dst[c] = self.buf[other].value + c * chunk_bytes
            lib.call('qb_permute_scatter', self.state._h, perm, k, dst, 0)
            self.state.sync()
            dev = f'cuda:{self.device}'
            nf = 2 << nl
            t_in = torch.as_tensor(_CudaArray(self.buf[other].value, nf), device=dev)
            t_out = torch.as_tensor(_CudaArray(self.buf[self.cur].value, nf), device=dev)
            splits = [0] * comm.world
            for c in range(1 << k):
                splits[peer_rank(c)] = nf >> k
            comm.all_to_all(t_out, t_in, splits, splits)
            torch.cuda.synchronize(self.device)
            self.exchange_seconds += time.perf_counter() - t0
            # the live shard is again buf[cur]
        self.exchanged_bytes += chunk_bytes * ((1 << k) - 1)
        self.exchanges += 1

    def probs_local(self, positions: Sequence[int]) -> np.ndarray:
        self.flush()
        out = np.empty(1 << len(positions), dtype=np.float64)
        self._lib.call('qb_probs', self.state._h, self._lib.int_array(positions), len(positions),
                       out.ctypes.data_as(C.c_void_p))
        return out

    def download_range(self, first: int, count: int) -> np.ndarray:
        self.flush()
        return self.state.download_range(first, count)

    def download(self) -> np.ndarray:
        self.flush()
        out = np.empty(1 << self.nl, dtype=np.complex128)
        self._lib.call('qb_download', self.state._h, out.ctypes.data_as(C.c_void_p), out.nbytes)
        return out

    def reduce_device(self):
        return f'cuda:{self.device}' if getattr(self.comm, 'dist', None) is not None and \
            self.comm.dist.get_backend(self.comm.group) == 'nccl' else None

    def close(self):
        lib = self._lib
        _trace("close: sync")
        self.sync()
        if self.peer:
            import torch
            torch.cuda.synchronize(self.device)
            _trace("close: barrier 1")
            self.comm.barrier()
            _trace("close: unmapping the peers")
            for r, ptrs in enumerate(self.peer):
                if r != self.comm.rank:
                    for p in ptrs + [self.peer_flags[r]]:
                        lib.call('qb_ipc_close', self.device, C.c_void_p(p))
            self.peer = None
            self.peer_flags = None
            _trace("close: barrier 2")
            self.comm.barrier()
        _trace("close: freeing")
        if self.flags.value:
            lib.call('qb_buffer_free', self.device, self.flags)
            self.flags.value = None
        self.state = None
        for b in self.buf:
            if b.value:
                lib.call('qb_buffer_free', self.device, b)
                b.value = None


# ---------------------------------------------------------------------------------------------
# the sharded register
# ---------------------------------------------------------------------------------------------
class ShardedKet:
    """n-qubit ket over comm.world = 2^g ranks.  Qubit arguments use the reference's numbering
    (qubit 0 = most significant index bit, qbot/qgates.py:161-182)."""

    def __init__(self, nq: int, comm, shard_factory=None, device: Optional[int] = None, exchange: str = 'p2p',
                 split: Optional[int] = None):
        world = comm.world
        # pieces (2^split) of a pipelined exchange.  Off by default: measured at 2 x B200 (33 qubits) it
        # does not pay -- the sweeps are HBM-bound, so the exchange's own HBM traffic cannot hide behind them, and what can
        # (the NVLink wait) is eaten by the SMs the peer stores need, or by the extra pack pass of the copy-engine variant
        self.split = int(os.environ.get('QBOT_B200_EXCHANGE_SPLIT', '0')) if split is None else int(split)
        self.min_first_phase = int(os.environ.get('QBOT_B200_EXCHANGE_SPLIT_MIN_GATES', '40'))   # fewer gates in the overlapped phases: not worth a split
        # gates per overlapped phase (before / after the exchange): about what it takes to cover the exchange (30 gates per
        # sweep, exchange = 2-3 sweeps); more would only keep more sweeps on the reduced grid
        self.phase_cap = int(os.environ.get('QBOT_B200_EXCHANGE_PHASE_GATES', '96'))
        g = world.bit_length() - 1
        if 1 << g != world:
            raise ValueError("the number of ranks must be a power of two")
        if nq - g < 1:
            raise ValueError("too few qubits for this many ranks")
        self.nq, self.comm, self.rank = nq, comm, comm.rank
        self.map = QubitMap(nq, g)
        if shard_factory is None:
            if device is None:
                raise ValueError("device is required for the CUDA shard")
            shard_factory = lambda nl, cm: CudaShard(nl, cm, device, exchange)     # noqa: E731
        self.shard = shard_factory(nq - g, comm)
        self.queue: List[LGate] = []
        self.gates_applied = 0
        self.label = list(range(nq))     # qubit of the caller -> qubit of the stored ket (`swap` only relabels)
        self._fresh: Optional[List[np.ndarray]] = None       # factors of a product ket that has not been written yet (init_product)
        self.lazy_map = os.environ.get('QBOT_B200_LAZY_MAP', '1') != '0'
        self.shard.init_basis(self.rank == 0, 0)

    def init_product(self, factors: Sequence[np.ndarray]):
        """The product ket kron(factors[0], factors[1], ...) (one 2-vector per qubit, qubit 0 first;
        density.tensorProd, qbot/density.py:7-24) with the identity qubit map: the rank bits are
        qubits 0 .. g-1, so this rank holds kron(local factors) times the product of the rank qubits'
        entries selected by its rank bits."""
        if len(factors) != self.nq:
            raise ValueError("one factor per qubit")
        self.queue = []
        self.map = QubitMap(self.nq, self.map.g)
        self.label = list(range(self.nq))
        # Nothing is written yet: the first flush picks the qubits that start on the rank bits from the gates queued
        # by then (QubitMap.choose_initial) and builds the shards for that map.
        self._fresh = [np.array(f, dtype=np.complex128).reshape(2) for f in factors]
        if not self.lazy_map:
            self._realise()

    def _realise(self):
        """Write the product ket of a pending init_product into the shards (called by every rank at the same point)."""
        factors = self._fresh
        if factors is None:
            return
        self._fresh = None
        self.shard.sync()
        self.comm.barrier()
        mp = self.map
        if self.lazy_map:
            mp.choose_initial(self.queue)
        n, nl, g = self.nq, mp.nl, mp.g
        of_bit = lambda b: factors[n - 1 - b]               # noqa: E731  logical index bit b <-> qubit n-1-b of the stored ket
        coeff = 1.0 + 0.0j
        for r in range(g):                                   # rank bit r = physical position nl + r
            coeff *= complex(of_bit(mp.at[nl + r])[(self.rank >> r) & 1])
        self.shard.init_product([of_bit(mp.at[p]) for p in range(nl - 1, -1, -1)], coeff)

    # -- gates ---------------------------------------------------------------------------------
    def _bit(self, q: int) -> int:
        return self.nq - 1 - self.label[int(q)]

    def swap_qubits(self, a: int, b: int):
        """genSwapGate (qbot/qgates.py:77-133) as a relabelling: later gates on qubit a act on what was b."""
        self.label[a], self.label[b] = self.label[b], self.label[a]
        return self

    def apply_gate_qubits(self, matrix, qubits: Sequence[int], controls: Iterable[int] = ()):
        """A gate on an arbitrary (not necessarily contiguous) list of qubits, qubits[0] = matrix MSB."""
        m = np.asarray(matrix)
        if m.shape != (1 << len(qubits), 1 << len(qubits)):
            raise ValueError("matrix does not match the number of qubits")
        self.queue.append(make_lgate(m, [self._bit(q) for q in qubits], [self._bit(c) for c in controls]))
        return self

    def apply_gate(self, matrix, first_target: int = 0, controls: Iterable[int] = ()):
        m = np.asarray(matrix)
        k = int(m.shape[0]).bit_length() - 1
        if first_target < 0 or first_target + k - 1 >= self.nq:
            raise IndexError(f"{k} qubit gate does not fit the {self.nq} qubit hilbertspace when started on qubit {first_target}")
        tb = [self._bit(first_target + j) for j in range(k)]
        self.queue.append(make_lgate(m, tb, [self._bit(c) for c in controls]))
        return self

    def flush(self):
        self._realise()
        rem = self.queue
        self.queue = []
        mp = self.map
        split = self.split if getattr(self.shard, 'supports_split', False) else 0
        while rem:
            picked = select_pass(rem, mp.local_mask())
            ps = set(picked)
            pass_gates = [rem[i] for i in picked]
            rem = [g for i, g in enumerate(rem) if i not in ps]
            for g in pass_gates:
                loc = mp.localise(g, self.rank)
                if loc is not None:
                    self.shard.apply(*loc)
            self.gates_applied += len(pass_gates)
            if not rem:
                break
            ex = mp.plan_exchange(rem, split, self.min_first_phase, prev=pass_gates if split else None, phase_cap=self.phase_cap)
            if ex.k == 0 and not picked:
                raise RuntimeError("sharded planner made no progress")
            if ex.split:
                # Runs the queued pass; those of its last sweeps and of the next pass's first sweeps that leave the parked
                # bits out of their tiles run sub-block by sub-block beside the pieces.  (Planning the gates that leave the
                # parked bits alone as lists of their own makes every sweep of them eligible -- 5.5 instead of 1 per step at
                # 8 GPUs -- but cuts a step into four plans: +2 sweeps per step, 160 ms against 152 serial.  Measured, not kept.)
                self.shard.do_exchange_split(ex)
            else:
                self.shard.flush()
                self.shard.do_exchange(ex)
        self.shard.flush()

    def sync(self):
        self.flush()
        self.shard.sync()

    def reset_zero(self):
        """|0...0> again, identity qubit map (a fresh register without re-allocating the shards)."""
        self.queue = []
        self._fresh = None
        self.shard.sync()
        self.comm.barrier()
        self.map = QubitMap(self.nq, self.map.g)
        self.label = list(range(self.nq))
        self.shard.init_basis(self.rank == 0, 0)

    # -- read-outs -----------------------------------------------------------------------------
    def probs(self, qubits: Sequence[int]) -> np.ndarray:
        """Outcome weights of the listed qubits (first listed = most significant outcome bit);
        local binning + one all-reduce of 2^m doubles."""
        self.flush()
        mp, nl = self.map, self.map.nl
        m = len(qubits)
        loc = [(i, mp.pos[self._bit(q)]) for i, q in enumerate(qubits) if mp.pos[self._bit(q)] < nl]
        glob = [(i, mp.pos[self._bit(q)] - nl) for i, q in enumerate(qubits) if mp.pos[self._bit(q)] >= nl]
        pl = self.shard.probs_local([p for _, p in loc])
        out = np.zeros(1 << m, dtype=np.float64)
        base = 0
        for i, r in glob:
            base |= ((self.rank >> r) & 1) << (m - 1 - i)
        ml = len(loc)
        for j in range(1 << ml):
            idx = base
            for x, (i, _) in enumerate(loc):
                idx |= ((j >> (ml - 1 - x)) & 1) << (m - 1 - i)
            out[idx] = pl[j]
        return self.comm.allreduce_sum(out, self.shard.reduce_device())

    def norm2(self) -> float:
        return float(self.probs([])[0])

    def amplitudes(self, indices: Sequence[int]) -> np.ndarray:
        """Amplitudes at the given basis-state indices (reference index convention)."""
        self.flush()
        mp, nl = self.map, self.map.nl
        out = np.zeros(2 * len(indices), dtype=np.float64)
        for x, idx in enumerate(indices):
            phys = 0
            for q in range(self.nq):
                phys |= ((idx >> (self.nq - 1 - q)) & 1) << mp.pos[self._bit(q)]
            if (phys >> nl) == self.rank:
                v = self.shard.download_range(phys & ((1 << nl) - 1), 1)[0]
                out[2 * x], out[2 * x + 1] = v.real, v.imag
        r = self.comm.allreduce_sum(out, self.shard.reduce_device())
        return r[0::2] + 1j * r[1::2]

    def gather(self) -> np.ndarray:
        """The full ket on every rank in the reference's index order (tests / small n only)."""
        self.flush()
        mp, n, nl = self.map, self.nq, self.map.nl
        local = self.shard.download()
        parts = self.comm.allgather_bytes(local.tobytes())
        phys = np.concatenate([np.frombuffer(p, dtype=np.complex128) for p in parts])
        # phys index bit p holds logical bit at[p]
        t = phys.reshape([2] * n)                      # axis a <-> physical bit n-1-a
        axes = [n - 1 - mp.pos[self._bit(q)] for q in range(n)]    # the caller's qubit q takes this physical axis
        return np.ascontiguousarray(t.transpose(axes)).reshape(-1)

    def make_local(self, qubits: Sequence[int]):
        """Exchange so that none of the listed qubits selects the rank (read-outs that need them local)."""
        self.flush()
        bits = [self._bit(q) for q in qubits]
        if all(self.map.is_local(b) for b in bits):
            return
        w = 0
        for b in bits:
            w |= 1 << b
        touch = LGate(np.eye(1 << len(bits), dtype=np.complex128), tuple(bits), (), w, 0)   # writes nothing, pins the bits
        ex = self.map.plan_exchange([touch])
        self.shard.flush()
        self.shard.do_exchange(ex)

    def reduced_density(self, qubits: Sequence[int]) -> np.ndarray:
        """Tr_rest psi psi^dagger on the listed qubits (first listed = most significant): the qubits are
        made local, every rank sums over its shard, one all-reduce adds the ranks."""
        if len(qubits) > 10:
            raise ValueError("reduced density of a sharded ket: at most 10 qubits")
        self.make_local(qubits)
        local = self.shard.rdm_local([self.map.pos[self._bit(q)] for q in qubits])
        d = 1 << len(qubits)
        flat = np.ascontiguousarray(np.asarray(local, dtype=np.complex128).reshape(-1)).view(np.float64)
        out = self.comm.allreduce_sum(flat, self.shard.reduce_device())
        return np.ascontiguousarray(out).view(np.complex128).reshape(d, d)

    def close(self):
        self.shard.close()


# ---------------------------------------------------------------------------------------------
# ProbVal branch batches sharded over ranks (BASELINE config 4)
# ---------------------------------------------------------------------------------------------
class ShardedBranchBatch:
    """B independent branch kets of n qubits, contiguous blocks of B/P branches per rank (keeps
    the funcWrapper ordering, qbot/probVal.py:365-375).  No data-path communication; the only
    collective is the final gather of the outcome weights."""

    def __init__(self, nq: int, nbranch: int, comm, device: int, factors: np.ndarray = None):
        from .state import DeviceState, KET
        if nbranch % comm.world:
            raise ValueError("the number of branches must be a multiple of the number of ranks")
        self.nq, self.nbranch, self.comm = nq, nbranch, comm
        self.per = nbranch // comm.world
        self.first = comm.rank * self.per
        self.device = device
        if factors is not None:
            self.state = DeviceState.product_batch(np.asarray(factors)[self.first:self.first + self.per], device)
        else:
            self.state = DeviceState.zero_state(nq, KET, self.per, device)

    def apply_gate(self, matrix, first_target: int = 0, controls: Iterable[int] = ()):
        """The same gate on every branch (fused engine, all branches in one sweep)."""
        self.state.apply_gate(matrix, first_target, controls)
        return self

    def apply_gate_per_branch(self, matrices, first_targets, controls=None, enable=None):
        """Branch b applies matrices[b] at first_targets[b] (global branch numbering)."""
        sl = slice(self.first, self.first + self.per)
        self.state.apply_gate_batched(np.asarray(matrices)[sl], list(first_targets)[sl],
                                      None if controls is None else list(controls)[sl],
                                      None if enable is None else list(enable)[sl])
        return self

    def probs(self, qubits: Sequence[int]) -> np.ndarray:
        """[nbranch, 2^m] on every rank: local outcome weights + one all-gather."""
        import torch
        local = np.ascontiguousarray(self.state.probs(qubits)).reshape(self.per, -1)
        dist = self.comm.dist
        backend = dist.get_backend(self.comm.group)
        dev = f'cuda:{self.device}' if backend == 'nccl' else 'cpu'
        t = torch.from_numpy(local).to(dev)
        out = torch.empty((self.nbranch, local.shape[1]), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(out, t, group=self.comm.group)
        return out.cpu().numpy()

    def sync(self):
        self.state.sync()
