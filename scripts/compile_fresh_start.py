"""NVRTC-compile (sm_100a, no GPU) every sweep of the fresh-start step of the multi-GPU e2e program (rc(n, 10) as a .qb
program on 8 / 4 / 2 ranks, every rank's local gate lists as the chosen start map plans them) and print registers / stack of
each cubin -- a build-container check that the kernels the round-end multi-GPU run will ask NVRTC for do compile, within
the 255-register budget and without local memory.

    python scripts/compile_fresh_start.py

TEST / ANALYSIS INFRASTRUCTURE."""
import sys, os, tempfile, time, subprocess, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'scripts')):
    sys.path.insert(0, p_)
import numpy as np
from plan_sharded_dry import OneRank, RecordingShard
from qbot_b200 import circuits, _lib
from qbot_b200.sharded import ShardedKet
tot = 0
for world, n in ((8,34),(4,34),(2,33)):
    gates = circuits.rc(n, 10, n)
    zero = [np.array([1,0],dtype=complex)]*n
    for rank in range(world):
        sk = ShardedKet(n, OneRank(rank, world), shard_factory=RecordingShard)
        sk.init_product(zero)
        for g in gates: sk.apply_gate(g.matrix(), g.target, g.controls)
        sk.flush()
        for seg in sk.shard.segments:
            if isinstance(seg, str): continue
            d = tempfile.mkdtemp(prefix='qb_fresh_')
            t0=time.perf_counter()
            k = _lib.jit_check(sk.map.nl, seg, d)
            regs = []
            for f in sorted(os.listdir(d)):
                if f.endswith('.cubin'):
                    out = subprocess.run(['cuobjdump','-res-usage',os.path.join(d,f)],capture_output=True,text=True).stdout
                    m = re.search(r'REG:(\d+).*?STACK:(\d+)', out)
                    regs.append((int(m.group(1)), int(m.group(2))) if m else None)
            tot += k
            print(f"world {world} rank {rank}: {len(seg)} gates -> {k} kernels compiled in {time.perf_counter()-t0:.1f}s (regs, stack): {regs}", flush=True)
print('total kernels', tot)
