"""ProbVal -- probabilistic values, same observable behaviour as the reference's
``qbot/probVal.py`` (normalise rules 22-51, flattening 53-70, constructors 72-96,
toDensityMatrix 99-111, operators 155-343, funcWrapper 347-390).

Quirks kept on purpose because programs can observe them:
  * ``normalize`` drops entries with p < 1e-5, *drops* (does not add) the probability of a
    later duplicate value, renormalises and rounds to 15 decimals;
  * a ProbVal (op) ProbVal binary operation evaluates ``op(other_value, self_value)`` when
    not reflected (probVal.py:183-193);
  * ``fan_out`` (funcWrapper) enumerates the Cartesian product with the FIRST ProbVal
    argument varying fastest.
"""
from __future__ import annotations

import math
import operator
from typing import Callable, List, Tuple

import numpy as np

smallVal = 1e-5
probRounding = 15


def _is_state(x) -> bool:
    return getattr(x, '_qb_device_state', False)


def valsClose(a, b) -> bool:
    if isinstance(a, float):
        return abs(a - b) < smallVal
    if _is_state(a) or _is_state(b):
        return bool(np.array_equal(np.asarray(a), np.asarray(b)))
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        return bool((a == b).all())
    return a == b


class ProbVal:
    __slots__ = ('probs', 'values')

    def __init__(self, probs, values):
        if len(probs) != len(values):
            raise Exception("len of probs and values must be the same")
        self.probs: List[float] = []
        self.values: list = []
        for p, v in zip(probs, values):
            if isinstance(v, ProbVal):            # nested ProbVals are flattened
                for sp, sv in zip(v.probs, v.values):
                    self.probs.append(p * sp)
                    self.values.append(sv)
            else:
                self.probs.append(p)
                self.values.append(v)
        self.normalize()

    @staticmethod
    def _exact_keys(values):
        """Hashable keys that are equal exactly when valsClose is true, for the value kinds whose
        equality is exact (same-shape ndarrays, gate descriptors, ints / bools / strings); None when
        the list holds anything else (floats compare with a tolerance, which is not an equivalence
        relation).  SURVEY.md row f4: the pairwise loop below is O(B^2) -- 8 million descriptor
        comparisons for a 4096-branch ProbVal."""
        if len(values) < 32:
            return None
        keys = []
        shape = None
        for x in values:
            if isinstance(x, np.ndarray):
                if x.dtype == object or (shape is not None and x.shape != shape):
                    return None
                shape = x.shape
                c = np.ascontiguousarray(x, dtype=complex) + 0.0          # -0.0 == 0.0 must hash alike
                if np.isnan(c.real).any() or np.isnan(c.imag).any():
                    return None
                keys.append(('a', x.shape, c.tobytes()))
            elif hasattr(x, 'canonical') and hasattr(x, 'controls'):      # GateDesc: equality of the full-space unitaries
                sup, m = x.canonical()
                m = np.ascontiguousarray(m, dtype=complex) + 0.0
                if np.isnan(m.real).any() or np.isnan(m.imag).any():
                    return None
                keys.append(('g', tuple(sup), m.shape, m.tobytes()))
            elif isinstance(x, (bool, int, str)) and not isinstance(x, float):
                keys.append(('s', x))
            else:
                return None
        kinds = {k[0] for k in keys}
        return keys if len(kinds) == 1 else None

    def normalize(self):
        p, v = self.probs, self.values
        keys = self._exact_keys(v)
        if keys is not None:
            # same outcome as the pairwise loop: an entry below smallVal is dropped when reached, a
            # kept entry removes every later equal one (whatever its probability)
            seen = set()
            np_, nv = [], []
            for pi, vi, ki in zip(p, v, keys):
                if pi < smallVal or ki in seen:
                    continue
                seen.add(ki)
                np_.append(pi)
                nv.append(vi)
            p[:] = np_
            v[:] = nv
            total = sum(p)
            for i in range(len(p)):
                p[i] = round(p[i] / total, probRounding)
            return
        i = 0
        while i < len(p):
            if p[i] < smallVal:
                del p[i], v[i]
                continue
            j = i + 1
            while j < len(p):
                if valsClose(v[i], v[j]):
                    del p[j], v[j]
                else:
                    j += 1
            i += 1
        total = sum(p)
        for i in range(len(p)):
            p[i] = round(p[i] / total, probRounding)

    # -- constructors that unwrap single-valued results ------------------------------------
    @staticmethod
    def fromUnzipped(probs, values):
        if len(values) == 1:
            return values[0]
        pv = ProbVal(probs, values)
        return pv.values[0] if len(pv.probs) == 1 else pv

    @staticmethod
    def fromZipped(pairs):
        if len(pairs) == 1:
            return pairs[0][1]
        pv = ProbVal([p for p, _ in pairs], [v for _, v in pairs])
        return pv.values[0] if len(pv.probs) == 1 else pv

    # -- inspection -----------------------------------------------------------------------------
    def instance(self):
        if not self.values:
            return None
        first = self.values[0]
        for v in self.values[1:]:
            if not isinstance(v, type(first)):
                return None
        return first

    def typeString(self):
        inst = self.instance()
        return "ProbVal<mixed>" if inst is None else f"ProbVal<{type(inst).__name__}>"

    def isEquivalent(self, other) -> bool:
        if not isinstance(other, ProbVal) or len(self.probs) != len(other.probs):
            return False
        for p, v in zip(self.probs, self.values):
            try:
                k = other.values.index(v)
            except ValueError:
                return False
            if not abs(p - other.probs[k]) < smallVal:
                return False
        return True

    def map(self, func):
        return ProbVal.fromUnzipped(self.probs, [func(v) for v in self.values])

    def toDensityMatrix(self):
        """sum_b p_b rho_b; 1-D values become psi psi^T first (probVal.py:99-111).
        Device-resident values are mixed on the device (ensemble kernel)."""
        inst = self.instance()
        if _is_state(inst):
            return type(inst).mix(self.probs, self.values)
        if isinstance(inst, np.ndarray):
            acc = np.zeros(self.values[0].shape, dtype=complex)
            for p, v in zip(self.probs, self.values):
                if v.ndim == 1:
                    v = np.outer(v, v)
                acc += p * v
            return acc
        raise TypeError()

    def __str__(self):
        return f"ProbVal({self.probs}, {self.values})"

    __hash__ = None

    # -- operator plumbing ------------------------------------------------------------------------
    def _unary(self, op, *args):
        return ProbVal.fromUnzipped(list(self.probs), [op(v, *args) for v in self.values])

    def _compare(self, other, op):
        yes = no = 0
        if isinstance(other, ProbVal):
            for p1, v1 in zip(self.probs, self.values):
                for p2, v2 in zip(other.probs, other.values):
                    if op(v1, v2):
                        yes += p1 * p2
                    else:
                        no += p1 * p2
        else:
            for p, v in zip(self.probs, self.values):
                if op(v, other):
                    yes += p
                else:
                    no += p
        return ProbVal.fromUnzipped([yes, no], [True, False])

    def _binary(self, other, op, reflected):
        probs, vals = [], []
        if isinstance(other, ProbVal):
            for p1, v1 in zip(self.probs, self.values):
                for p2, v2 in zip(other.probs, other.values):
                    vals.append(op(v1, v2) if reflected else op(v2, v1))   # sic, see module docstring
                    probs.append(p1 * p2)
        else:
            for p, v in zip(self.probs, self.values):
                probs.append(p)
                vals.append(op(other, v) if reflected else op(v, other))
        return ProbVal.fromUnzipped(probs, vals)

    def __round__(self, ndigits=None):
        return self._unary(round, ndigits)

    def __not__(self):
        return self._unary(operator.not_)

    def __div__(self, other):
        return self.__truediv__(other)

    def __rdiv__(self, other):
        return self.__rtruediv__(other)


def _install_operators():
    for name, op in (('eq', operator.eq), ('ne', operator.ne), ('gt', operator.gt), ('lt', operator.lt),
                     ('ge', operator.ge), ('le', operator.le)):
        setattr(ProbVal, f'__{name}__', (lambda o: lambda self, other: self._compare(other, o))(op))
    # the reference routes the logical operators through the comparison path (truthiness of op result)
    for name, op in (('and', operator.and_), ('or', operator.or_), ('xor', operator.xor)):
        f = (lambda o: lambda self, other: self._compare(other, o))(op)
        setattr(ProbVal, f'__{name}__', f)
        setattr(ProbVal, f'__r{name}__', f)
    for name, op in (('abs', operator.abs), ('trunc', math.trunc), ('floor', math.floor), ('ceil', math.ceil),
                     ('neg', operator.neg), ('invert', operator.inv), ('pos', operator.pos)):
        setattr(ProbVal, f'__{name}__', (lambda o: lambda self: self._unary(o))(op))
    for name, op in (('add', operator.add), ('sub', operator.sub), ('mul', operator.mul),
                     ('truediv', operator.truediv), ('mod', operator.mod), ('floordiv', operator.floordiv),
                     ('lshift', operator.lshift), ('rshift', operator.rshift), ('matmul', operator.matmul)):
        setattr(ProbVal, f'__{name}__', (lambda o: lambda self, other: self._binary(other, o, False))(op))
        setattr(ProbVal, f'__r{name}__', (lambda o: lambda self, other: self._binary(other, o, True))(op))


_install_operators()


def funcWrapper(func: Callable, *args, **kwargs):
    """Make ``func`` probabilistic: call it once per combination of the ProbVal arguments'
    branches (first ProbVal argument fastest) and collect ``ProbVal.fromUnzipped``."""
    lens = [len(a.probs) for a in args if isinstance(a, ProbVal)]
    lens += [len(v.probs) for v in kwargs.values() if isinstance(v, ProbVal)]
    if not lens:
        return func(*args, **kwargs)          # one certain branch: fromUnzipped([1], [v]) is v (probVal.py:89-95)
    total = 1
    for n in lens:
        total *= n
    probs, vals = [], []
    for it in range(total):
        rem, weight = it, 1
        call_args = []
        for a in args:
            if isinstance(a, ProbVal):
                k = rem % len(a.probs)
                rem //= len(a.probs)
                weight *= a.probs[k]
                call_args.append(a.values[k])
            else:
                call_args.append(a)
        call_kwargs = {}
        for key, v in kwargs.items():
            if isinstance(v, ProbVal):
                k = rem % len(v.probs)
                rem //= len(v.probs)
                weight *= v.probs[k]
                call_kwargs[key] = v.values[k]
            else:
                call_kwargs[key] = v
        probs.append(weight)
        vals.append(func(*call_args, **call_kwargs))
    return ProbVal.fromUnzipped(probs, vals)
