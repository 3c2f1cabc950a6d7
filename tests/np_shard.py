"""Test-only shard backends for the sharded-ket host logic: a numpy shard (no GPU) whose
exchange goes either through torch.distributed (gloo) or through an in-process communicator
that runs P virtual ranks as threads."""
import threading

import numpy as np


def np_apply_bits(psi, nbits, m, tbits, cmask):
    """psi <- gate m on index bits tbits (tbits[0] = matrix MSB) where all cmask bits are 1."""
    k = len(tbits)
    idx = np.arange(1 << nbits, dtype=np.int64)
    sel = (idx & cmask) == cmask
    row = np.zeros(1 << nbits, dtype=np.int64)
    for j, b in enumerate(tbits):
        row |= ((idx >> b) & 1) << (k - 1 - j)
    tm = 0
    for b in tbits:
        tm |= 1 << b
    base = idx & ~tm
    out = psi.copy()
    acc = np.zeros(1 << nbits, dtype=complex)
    for col in range(1 << k):
        off = 0
        for j, b in enumerate(tbits):
            off |= ((col >> (k - 1 - j)) & 1) << b
        acc += m[row, col] * psi[base | off]
    out[sel] = acc[sel]
    return out


class VirtualComm:
    """P ranks as threads of one process."""

    class Shared:
        def __init__(self, world):
            self.world = world
            self.barrier = threading.Barrier(world)
            self.slots = [None] * world

    def __init__(self, shared, rank):
        self.sh, self.rank, self.world = shared, rank, shared.world
        self.dist = None
        self.group = None

    def barrier(self):
        self.sh.barrier.wait()

    def _exchange(self, payload):
        self.sh.slots[self.rank] = payload
        self.sh.barrier.wait()
        got = list(self.sh.slots)
        self.sh.barrier.wait()
        return got

    def allreduce_sum(self, arr, device=None):
        return np.sum(self._exchange(np.array(arr, dtype=np.float64)), axis=0)

    def allgather_bytes(self, payload):
        return self._exchange(payload)

    def all_to_all_np(self, chunks_by_rank):
        """chunks_by_rank[r] = array for rank r (or None); returns what every rank sent to me."""
        got = self._exchange(chunks_by_rank)
        return [got[r][self.rank] for r in range(self.world)]


class NumpyShard:
    def __init__(self, nl, comm):
        self.nl, self.comm = nl, comm
        self.psi = np.zeros(1 << nl, dtype=complex)
        self.exchanges = 0
        self.exchanged_bytes = 0
        self.applied = 0

    def init_basis(self, has_one, local_index=0):
        self.psi[:] = 0
        if has_one:
            self.psi[local_index] = 1

    def init_product(self, local_factors, coeff):
        v = np.array([coeff], dtype=complex)
        for f in local_factors:
            v = np.kron(v, np.asarray(f, dtype=complex).reshape(2))
        self.psi = v

    def rdm_local(self, positions):
        nl = self.nl
        t = self.psi.reshape([2] * nl)                       # axis a <-> local bit nl-1-a
        keep = [nl - 1 - p for p in positions]
        rest = [a for a in range(nl) if a not in keep]
        m = np.ascontiguousarray(t.transpose(keep + rest)).reshape(1 << len(keep), -1)
        return m @ m.conj().T

    def apply(self, m, tpos, cmask):
        self.psi = np_apply_bits(self.psi, self.nl, np.asarray(m), list(tpos), cmask)
        self.applied += 1

    def flush(self):
        pass

    def sync(self):
        pass

    def do_exchange(self, ex):
        k, nl = ex.k, self.nl
        if k == 0:
            return
        j = np.arange(1 << nl, dtype=np.int64)
        src = np.zeros_like(j)
        for d, f in enumerate(ex.src_bit_of_dst_bit):
            src |= ((j >> d) & 1) << f
        packed = self.psi[src]
        rank = self.comm.rank

        def peer_rank(c):
            pr = rank
            for i, r in enumerate(ex.rank_bits):
                pr = (pr & ~(1 << r)) | (((c >> i) & 1) << r)
            return pr
        # packed layout, top down: [v parked bits][k chunk bits][rest] (v = ex.split; 0 = the plain exchange)
        v = getattr(ex, 'split', 0)
        chunks = packed.reshape(1 << v, 1 << k, -1)
        if getattr(self.comm, 'dist', None) is not None:
            import torch
            send = np.ascontiguousarray(chunks.transpose(1, 0, 2))         # [chunk][piece][rest]
            t_in = torch.from_numpy(send.reshape(-1).view(np.float64))
            t_out = torch.zeros_like(t_in)
            splits = [0] * self.comm.world
            for c in range(1 << k):
                splits[peer_rank(c)] = t_in.numel() >> k
            self.comm.all_to_all(t_out, t_in, splits, splits)
            # arrives ordered by source RANK; chunk s of the new shard comes from peer_rank(s)
            got = t_out.numpy().view(np.complex128).reshape(1 << k, 1 << v, -1)
            by_rank = sorted(range(1 << k), key=peer_rank)
            new = np.empty_like(chunks)
            for slot, s_ in enumerate(by_rank):
                new[:, s_, :] = got[slot]
            self.psi = new.reshape(-1).copy()
        else:
            send = [None] * self.comm.world
            for c in range(1 << k):
                send[peer_rank(c)] = chunks[:, c, :].copy()
            got = self.comm.all_to_all_np(send)
            new = np.empty_like(chunks)
            for s_ in range(1 << k):
                new[:, s_, :] = got[peer_rank(s_)]
            self.psi = new.reshape(-1)
        self.exchanges += 1
        self.exchanged_bytes += (16 << nl) * ((1 << k) - 1) >> k

    supports_split = True      # (tests switch it off to compare with the plain exchange)

    def do_exchange_split(self, ex):
        """The data movement of a pipelined exchange (packed layout [parked bits][chunk bits][rest]) in one go; what
        CudaShard overlaps with it -- sweeps on tile ranges -- does not exist on the numpy shard."""
        self.do_exchange(ex)
        self.split_exchanges = getattr(self, 'split_exchanges', 0) + 1
        self.send_side_exchanges = getattr(self, 'send_side_exchanges', 0) + (1 if ex.tail else 0)

    def probs_local(self, positions):
        m = len(positions)
        idx = np.arange(1 << self.nl, dtype=np.int64)
        key = np.zeros_like(idx)
        for x, p in enumerate(positions):
            key |= ((idx >> p) & 1) << (m - 1 - x)
        return np.bincount(key, weights=np.abs(self.psi) ** 2, minlength=1 << m)

    def download_range(self, first, count):
        return self.psi[first:first + count].copy()

    def download(self):
        return self.psi.copy()

    def reduce_device(self):
        return None

    def close(self):
        pass
