"""Diagnostic (not part of the product): where a sharded-ket step spends its time.
torchrun --nproc-per-node N scripts/diag_sharded.py [qubits] [depth]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from qbot_b200 import circuits
from qbot_b200.sharded import ShardedKet, TorchComm, CudaShard

rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); local = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 33
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 10
gates = circuits.rc(n, depth, n)
mats = [np.ascontiguousarray(g.matrix()) for g in gates]
sk = ShardedKet(n, TorchComm(), device=local)
log = []
orig_flush = CudaShard.flush
def timed_flush(self):
    st = self.state
    t0 = time.perf_counter()
    s0 = st.stats()
    st.timer_start()
    orig_flush(self)
    ms = st.timer_stop()
    s1 = st.stats()
    log.append(('flush', ms, (time.perf_counter() - t0) * 1e3, s1['fused_passes'] - s0['fused_passes'], s1['jit_passes'] - s0['jit_passes'],
                s1['kernel_launches'] - s0['kernel_launches']))
CudaShard.flush = timed_flush
orig_ex = CudaShard.do_exchange
def timed_ex(self, ex):
    t0 = time.perf_counter()
    e0 = self.exchange_seconds
    orig_ex(self, ex)
    log.append(('exchange', (self.exchange_seconds - e0) * 1e3, (time.perf_counter() - t0) * 1e3, ex.k, 0, 0))
CudaShard.do_exchange = timed_ex
for step in range(7):
    log.clear()
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for g, m in zip(gates, mats):
        sk.apply_gate(m, g.target, g.controls)
    t1 = time.perf_counter()
    sk.flush(); sk.shard.sync()
    t2 = time.perf_counter()
    if rank == 0:
        print(f"step {step}: queue {1e3*(t1-t0):.1f} ms, flush {1e3*(t2-t1):.1f} ms")
        for e in log:
            print("   ", e[0], f"gpu/ex {e[1]:.1f} ms wall {e[2]:.1f} ms", "sweeps/k", e[3], "jit", e[4], "launches", e[5])
sk.close()
dist.barrier()
dist.destroy_process_group()
