"""qbot_b200 -- B200-native state-path backend for the qbot DSL.

Public surface
    executeTxt(text) / executeFile(fileobj)   run a qbot program on the CUDA backend; returns the
                                              final namespace dict like the reference's
                                              qbot.executeTxt / executeFile (interpreter.py:231-235)
    install() / uninstall()                   re-register the six state ops inside a real qbot
                                              checkout (qbot.operators.operations) -- see INTEGRATION.md
    DeviceState                               device-resident ket / density matrix / branch batch
    circuits.rc(n, depth, seed)               the benchmark circuit generator

There is no CPU fallback: the CUDA library (qbot_b200/lib/libqbot_b200.so, built by
``__graft_entry__.build()``) must be present and a CUDA device visible.
"""
__version__ = "0.1.0"

from .state import DeviceState, KET, DM          # noqa: F401
from . import circuits                           # noqa: F401


def _interpreter(state_cls=None):
    from .host.interp import Interpreter
    return Interpreter(state_cls or DeviceState)


def executeTxt(text: str, state_cls=None):
    return _interpreter(state_cls).execute(text.splitlines())


def executeFile(file, state_cls=None):
    return _interpreter(state_cls).execute(file.readlines())


def install(qbot_module=None):
    from .integration import install as _install
    return _install(qbot_module)


def uninstall():
    from .integration import uninstall as _uninstall
    return _uninstall()
