"""Expression namespace of the DSL (what ``qbot/evaluation.py`` exposes to ``eval``):
ProbVal constructors, gate constants, gate/tensor functions, basis aliases, and
``math_*`` / ``np_*`` / ``linalg_*`` wrappers that fan out over ProbVal arguments.
Expressions are evaluated with ``__builtins__ = {}`` and the interpreter's local namespace,
so ``state`` is visible to user expressions."""
from __future__ import annotations

import math

import numpy as np

from . import hostmath as hm
from . import errors as err
from .probval import ProbVal, funcWrapper

_s = 2 ** (-1 / 2)


def _pv(func):
    return lambda *a, **k: funcWrapper(func, *a, **k)


def build_namespace():
    ns = {
        '__builtins__': {},
        'ProbVal': ProbVal.fromUnzipped,
        'ProbValZipped': ProbVal.fromZipped,
        'identityGate': np.eye(2),
        'hadamardGate': _s * np.array([[1, 1], [1, -1]], dtype=complex),
        'pauliXGate': np.array([[0, 1], [1, 0]], dtype=complex),
        'pauliYGate': np.array([[0, -1j], [1j, -0]], dtype=complex),
        'pauliZGate': np.array([[1, 0], [0, -1]], dtype=complex),
        'xRotGate': lambda theta: funcWrapper(hm.x_rot, theta),
        'yRotGate': lambda theta: funcWrapper(hm.y_rot, theta),
        'zRotGate': lambda theta: funcWrapper(hm.z_rot, theta),
        'qftGate': hm.qft,
        'simonsGate': lambda numQubits, f: funcWrapper(hm.simons, numQubits, f),
        'swapGate': lambda numQubits, a, b: funcWrapper(hm.swap_matrix, numQubits, a, b),
        'shiftGate': lambda numQubits, up=True, numShifts=1: funcWrapper(hm.shift_matrix, numQubits, up, numShifts),
        'plist': lambda *a: funcWrapper(lambda *x: list(x), *a),
        'ptuple': lambda *a: funcWrapper(lambda *x: tuple(x), *a),
        'pset': lambda *a: funcWrapper(lambda *x: set(x), *a),
        'tensorProd': _pv(hm.tensor_prod),
        'tensorExp': _pv(hm.tensor_exp),
        'tensorPermute': _pv(hm.tensor_permute),
        'ketToDensity': _pv(hm.ket_to_density),
        'ketsToDensity': _pv(hm.kets_to_density_zipped),
        'densityToKets': hm.density_to_kets,
    }
    for name in dir(math):
        obj = getattr(math, name)
        if name.startswith('_'):
            continue
        ns[f'math_{name}'] = _pv(obj) if callable(obj) else obj
    for name in dir(np):
        if name.startswith('_'):
            continue
        try:
            obj = getattr(np, name)
        except Exception:
            continue
        if callable(obj) and not isinstance(obj, type(np)):
            ns[f'np_{name}'] = _pv(obj)
    for name in dir(np.linalg):
        obj = getattr(np.linalg, name)
        if not name.startswith('_') and callable(obj) and not isinstance(obj, type):
            ns[f'linalg_{name}'] = _pv(obj)
    for b in hm.all_bases:
        for alias in b.names:
            ns[alias] = b
    return ns


globalNameSpace = build_namespace()


_code_cache = {}


def evaluate(expression: str, localNameSpace: dict):
    code = _code_cache.get(expression)
    if code is None:
        code = compile(expression, "<string>", "eval")
        if len(_code_cache) < 65536:
            _code_cache[expression] = code       # the same source lines are evaluated again by loops and re-runs
    return eval(code, globalNameSpace, localNameSpace)


def evaluateWrapper(lines, lineNum, expression: str, localNameSpace: dict):
    try:
        return evaluate(expression, localNameSpace)
    except Exception as e:
        err.raiseFormattedError(err.pythonError(lines, lineNum, e))
