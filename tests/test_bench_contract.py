"""CPU: the parts of bench.py's contract that need no GPU -- the reference arm (the reference's
algorithm timed on the host cores) prints ONE JSON line with the agreed keys, and non-zero ranks
of a torchrun launch stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_ref(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
                        '--ref-qubits', '7'], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout.strip()


def test_reference_arm_line():
    out = run_ref()
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'gates/sec' and d['unit'] == 'gates/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['steps'] == 1 and d['warmup'] == 0
    cb = d['cpu_baseline']
    assert cb['kind'] in ('port', 'reference') and cb['cores'] >= 1 and cb['value'] == d['value'] and 'sample' in cb
    e = d['e2e']
    assert e['value'] == d['value'] and e['h2d_bytes_per_step'] == 0 and e['d2h_bytes_per_step'] == 0
    assert 'workload' in d['config']


def test_reference_arm_other_ranks_are_silent():
    assert run_ref({'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'}) == ''


def test_traffic_capture_is_of_heads_kernel_set(tmp_path):
    """bench.py prints roofline.traffic only when profiles/r02_traffic.json holds an ncu capture of the
    kernel set it ran (looked up by qb_stats.jit_kernel_hash = the sum of the FNV-1a hashes of the
    generated sweep sources of one step).  The sources come out of the planner + specialiser, which need
    no GPU: regenerate the headline's sweeps here and check that their hash is one of the captured sets,
    i.e. that nobody changed the generator after the last capture without re-measuring."""
    sys.path.insert(0, ROOT)
    import plan_emu
    from qbot_b200 import _lib
    from qbot_b200.circuits import rc
    n = 30
    nk = _lib.jit_check(n, plan_emu.circuit_to_bits(n, rc(n, 20, 30)), str(tmp_path))
    total = 0
    srcs = sorted(f for f in os.listdir(tmp_path) if f.endswith('.cu'))
    assert len(srcs) == nk and nk > 0
    for f in srcs:
        h = 1469598103934665603
        for c in open(os.path.join(tmp_path, f), 'rb').read():
            h = ((h ^ c) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        total = (total + h) & 0xFFFFFFFFFFFFFFFF
    captured = json.load(open(os.path.join(ROOT, 'profiles', 'r02_traffic.json')))['qj_kernel']
    key = f"{total:016x}"
    assert key in captured, f"HEAD's kernel set {key} has no ncu capture (captured: {sorted(captured)})"
    assert captured[key]['qubits'] == n and captured[key]['launches_captured'] == nk
    # no wasted re-reads: measured DRAM traffic per launch within 1 % of the algorithmic 32 * 2^n bytes
    assert abs(captured[key]['dram_bytes_per_launch'] / (32 * 2 ** n) - 1) < 0.01
