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
