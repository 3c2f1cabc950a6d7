"""CPU: the C-ABI library loads and exports exactly what include/qbot_b200.h declares; the
ctypes prototype table covers every declared function; compute calls fail loudly without a
GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import __graft_entry__ as entry
from qbot_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, 'include', 'qbot_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(qb_[a-z0-9_]+)\s*\(', text)))


def test_library_builds_and_exports_header_symbols():
    entry.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.PROTOTYPES) == names, "ctypes prototype table and header differ"
    _lib.load()
    assert _lib.load().qb_version().decode().startswith('qbot_b200')


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from qbot_b200 import DeviceState
    with pytest.raises(_lib.QbotB200Error):
        DeviceState.zero_state(3)
    with pytest.raises(_lib.QbotB200Error):
        import qbot_b200
        qbot_b200.executeTxt("qset comp[0]\ngate hadamardGate ; 0\n")


def test_product_does_not_import_oracle():
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); import qbot_b200, qbot_b200.host.interp, qbot_b200.integration, qbot_b200.circuits; "
            "bad = [m for m in sys.modules if m.split('.')[0] == 'oracle']; assert not bad, bad") % ROOT
    assert subprocess.run([sys.executable, '-c', code]).returncode == 0
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'qbot_b200')):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in src and 'from oracle' not in src, f
