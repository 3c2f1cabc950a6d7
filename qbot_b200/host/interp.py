"""Line interpreter of the DSL: tokeniser, mark pre-pass, fetch / dispatch / jump loop, and the
non-state operations (defines, control flow, console).  Behaviour follows the reference's
``qbot/interpreter.py:82-235`` and the control-flow half of ``qbot/operators.py`` (191-252,
431-472); the state operations come from ``qbot_b200.host.ops``.

A program line is ``OPER arg1 ; arg2 ; ...``: the first four characters (case-insensitive)
name the operation, the remainder is split on ``;``.  ``note`` lines are comments and ``mark``
lines are jump targets recorded before execution starts.
"""
from __future__ import annotations

import numpy as np

from . import errors as err
from .namespace import evaluateWrapper
from .probval import ProbVal, funcWrapper
from . import hostmath as hm
from .ops import Host, make_ops, is_state


class OpReturnVal:
    def __init__(self, jumpLineNum=None, joinLineNum=None, halt=False):
        self.jumpLineNum = jumpLineNum
        self.joinLineNum = joinLineNum
        self.halt = halt


_token_cache = {}


def processLineIntoTokens(line: str):
    cached = _token_cache.get(line)
    if cached is not None:
        return list(cached)          # loops and re-runs meet the same source lines again
    text = line.strip()
    if not text:
        return []
    tokens = [text[:4].lower()]
    tokens += [s.strip() for s in text[4:].split(';') if s.strip()]
    if len(_token_cache) < 65536:
        _token_cache[line] = tuple(tokens)
    return tokens


def _var_name(lines, lineNum, token):
    if not token.isidentifier():
        err.raiseFormattedError(err.customInvalidVariableName(lines, lineNum, token))
    return token


def _mark_line(ns, lines, lineNum, token) -> int:
    if token.isidentifier() and token in ns['__marks']:
        return ns['__marks'][token]
    res = evaluateWrapper(lines, lineNum, token, ns)
    if isinstance(res, str):
        try:
            return ns['__marks'][res]
        except KeyError:
            err.raiseFormattedError(err.customUnknownMarkName(lines, lineNum, token))
    got = res.typeString() if isinstance(res, ProbVal) else type(res).__name__
    err.raiseFormattedError(err.customTypeError(lines, lineNum, ['str'], got))


def _bool_or_die(lines, lineNum, val):
    if isinstance(val, bool):
        return val
    got = val.typeString() if isinstance(val, ProbVal) else type(val).__name__
    err.raiseFormattedError(err.customTypeError(lines, lineNum, ['bool'], got))


class Interpreter:
    """One interpreter = one register (``localNameSpace['state']``), as in the reference."""

    def __init__(self, state_cls):
        host = Host(ProbVal, funcWrapper, evaluateWrapper, err, hm.Basis, state_cls)
        self.state_ops = make_ops(host)
        so = self.state_ops
        self.operations = {
            'cdef': (self.cdef, 2, 2), 'qdef': (self.qdef, 2, 2),
            'qset': (so['qset'], 1, 2), 'gate': (so['gate'], 1, 4), 'disc': (so['disc'], 1, 1), 'swap': (so['swap'], 2, 2),
            'meas': (so['meas'], 2, 3), 'peek': (so['peek'], 2, 3),
            'jump': (self.jump, 1, 1), 'cjmp': (self.cjmp, 2, 2), 'halt': (self.halt, 0, 1), 'retr': (self.retr, 0, 1),
            'pydo': (self.pydo, 1, 1), 'cout': (self.cout, 1, 1),
        }

    # ---- non-state operations ---------------------------------------------------------------------
    def _set(self, ns, key, value, quantum):
        ns[key] = value
        ns[f'__is_q_{key}'] = quantum
        ns[f'__updated_{key}'] = True

    def cdef(self, ns, lines, lineNum, tokens):
        name = _var_name(lines, lineNum, tokens[1])
        self._set(ns, name, evaluateWrapper(lines, lineNum, tokens[2], ns), False)

    def qdef(self, ns, lines, lineNum, tokens):
        name = _var_name(lines, lineNum, tokens[1])
        val = self.state_ops['convertToDensity'](lines, lineNum, evaluateWrapper(lines, lineNum, tokens[2], ns))
        self._set(ns, name, val, True)

    def jump(self, ns, lines, lineNum, tokens):
        ns['__prev_jump'] = lineNum
        return OpReturnVal(_mark_line(ns, lines, lineNum, tokens[1]))

    def cjmp(self, ns, lines, lineNum, tokens):
        target = _mark_line(ns, lines, lineNum, tokens[1])
        if _bool_or_die(lines, lineNum, evaluateWrapper(lines, lineNum, tokens[2], ns)):
            ns['__prev_jump'] = lineNum
            return OpReturnVal(target)

    def halt(self, ns, lines, lineNum, tokens):
        if len(tokens) < 2:
            return OpReturnVal(halt=True)
        return OpReturnVal(halt=_bool_or_die(lines, lineNum, evaluateWrapper(lines, lineNum, tokens[1], ns)))

    def retr(self, ns, lines, lineNum, tokens):
        if len(tokens) < 2 or _bool_or_die(lines, lineNum, evaluateWrapper(lines, lineNum, tokens[1], ns)):
            return OpReturnVal(jumpLineNum=ns['__prev_jump'] + 1)

    def pydo(self, ns, lines, lineNum, tokens):
        evaluateWrapper(lines, lineNum, tokens[1], ns)

    def cout(self, ns, lines, lineNum, tokens):
        print(evaluateWrapper(lines, lineNum, tokens[1], ns))

    # ---- driver -----------------------------------------------------------------------------------
    def record_marks(self, ns, lines):
        for lineNum, line in enumerate(lines):
            if line.strip()[:4].lower() == 'mark':
                tokens = processLineIntoTokens(line)
                name = tokens[1]
                if not name.isidentifier():
                    err.raiseFormattedError(err.customInvalidMarkName(lines, lineNum, name))
                ns['__marks'][name] = lineNum

    def runtime(self, ns, lines):
        lineNum = -1
        while lineNum < len(lines) - 1:
            lineNum += 1
            tokens = processLineIntoTokens(lines[lineNum])
            if not tokens or tokens[0] in ('note', 'mark'):
                continue
            try:
                op, lo, hi = self.operations[tokens[0]]
            except KeyError:
                err.raiseFormattedError(err.customUnknownOperationError(lines, lineNum, tokens[0]))
            nargs = len(tokens) - 1
            if nargs < lo or nargs > hi:
                err.raiseFormattedError(err.customNumArgumentsError(lines, lineNum, tokens[0], nargs, lo, hi))
            ret = op(ns, lines, lineNum, tokens)
            if ret is None:
                continue
            if isinstance(ret.halt, bool) and ret.halt:
                break
            if isinstance(ret.jumpLineNum, int):
                lineNum = ret.jumpLineNum - 1

    def execute(self, lines):
        ns = {'state': np.array([], dtype=complex), '__updated_state': False, '__marks': dict(), '__prev_jump': -1}
        self.record_marks(ns, lines)
        self.runtime(ns, lines)
        return ns
