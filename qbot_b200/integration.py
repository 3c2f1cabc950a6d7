"""Drop-in installation into a real qbot checkout.

``install()`` re-registers the six state operations in the reference's dispatch table
(``qbot/operators.py:477-506``; the interpreter holds the same dict object,
``qbot/interpreter.py:5,132``) so that an unchanged ``qbot.executeTxt`` / ``executeFile`` / CLI
runs its register on the GPU.  Two helpers of ``qbot.probVal`` learn about device-resident
values (exact-equality de-duplication and the ensemble sum); nothing else is touched.
"""
from __future__ import annotations

import numpy as np

_saved = None


def install(qbot_module=None, state_cls=None):
    global _saved
    if _saved is not None:
        return
    import importlib
    if qbot_module is None:
        qbot_module = importlib.import_module('qbot')
    ops_mod = importlib.import_module('qbot.operators')
    pv_mod = importlib.import_module('qbot.probVal')
    ev_mod = importlib.import_module('qbot.evaluation')
    err_mod = importlib.import_module('qbot.errors')
    basis_mod = importlib.import_module('qbot.basis')
    meas_mod = importlib.import_module('qbot.measurement')
    from .host.ops import Host, make_ops, is_state
    if state_cls is None:
        from .state import DeviceState as state_cls
    host = Host(pv_mod.ProbVal, pv_mod.funcWrapper, ev_mod.evaluateWrapper, err_mod, basis_mod.Basis, state_cls,
                MeasurementResult=meas_mod.MeasurementResult)
    new = make_ops(host)
    table = ops_mod.operations
    gns = getattr(ev_mod, 'globalNameSpace', None)
    _saved = dict(table={k: table[k] for k in ('qset', 'gate', 'disc', 'swap', 'meas', 'peek')},
                  valsClose=pv_mod.valsClose, toDensityMatrix=pv_mod.ProbVal.toDensityMatrix,
                  convert=ops_mod.convertToDensity, mods=(ops_mod, pv_mod), gns=gns, normalize=pv_mod.ProbVal.normalize,
                  tensor={k: gns[k] for k in ('tensorProd', 'tensorExp') if gns is not None and k in gns})
    if _saved['tensor']:
        # state constructors of >= 14 qubits stay descriptors and are built on the device
        # (SURVEY.md row f1); below that they are the reference's host kron chains
        from .host import hostmath as hm
        fw = pv_mod.funcWrapper
        gns['tensorProd'] = lambda *a, **k: fw(hm.tensor_prod, *a, **k)
        gns['tensorExp'] = lambda *a, **k: fw(hm.tensor_exp, *a, **k)
    for name in _saved['table']:
        _, lo, hi = table[name]
        table[name] = (new[name], lo, hi)

    ref_close, ref_tdm, ref_convert = _saved['valsClose'], _saved['toDensityMatrix'], _saved['convert']

    def valsClose(a, b):
        if is_state(a) or is_state(b):
            return bool(np.array_equal(np.asarray(a), np.asarray(b)))
        return ref_close(a, b)

    def toDensityMatrix(self):
        inst = self.instance()
        if is_state(inst):
            return type(inst).mix(self.probs, self.values)
        return ref_tdm(self)

    def convertToDensity(lines, lineNum, val):      # used by the reference's qdef
        return val if is_state(val) else ref_convert(lines, lineNum, val)

    ref_normalize = _saved['normalize']
    from .host.probval import ProbVal as _MirrorProbVal

    def normalize(self):
        """the reference's de-duplication is a pairwise loop, O(B^2) value comparisons
        (qbot/probVal.py:22-51) -- 8 million gate-descriptor comparisons for the 4096-branch ProbVal of
        BASELINE config 4.  For value kinds whose equality is exact (same-shape arrays, gate descriptors,
        ints / bools / strings) a hash set keeps exactly what that loop keeps (first occurrence wins, later
        duplicates are dropped without adding their probability, entries below smallVal go when reached);
        everything else -- floats compare with a tolerance -- takes the reference's own loop."""
        keys = _MirrorProbVal._exact_keys(self.values)
        if keys is None:
            return ref_normalize(self)
        seen, np_, nv = set(), [], []
        for pi, vi, ki in zip(self.probs, self.values, keys):
            if pi < pv_mod.smallVal or ki in seen:
                continue
            seen.add(ki)
            np_.append(pi)
            nv.append(vi)
        self.probs[:] = np_
        self.values[:] = nv
        total = sum(self.probs)
        for i in range(len(self.probs)):
            self.probs[i] /= total
            self.probs[i] = round(self.probs[i], pv_mod.probRounding)

    pv_mod.valsClose = valsClose
    pv_mod.ProbVal.toDensityMatrix = toDensityMatrix
    pv_mod.ProbVal.normalize = normalize
    ops_mod.convertToDensity = convertToDensity


def uninstall():
    global _saved
    if _saved is None:
        return
    ops_mod, pv_mod = _saved['mods']
    for name, entry in _saved['table'].items():
        ops_mod.operations[name] = entry
    for k, v in _saved.get('tensor', {}).items():
        _saved['gns'][k] = v
    pv_mod.valsClose = _saved['valsClose']
    pv_mod.ProbVal.toDensityMatrix = _saved['toDensityMatrix']
    pv_mod.ProbVal.normalize = _saved['normalize']
    ops_mod.convertToDensity = _saved['convert']
    _saved = None
