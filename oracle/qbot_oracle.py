"""CPU oracle for the qbot state path -- TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the arithmetic the reference (PrinceOfPuppers/qbot)
performs on its register: full-space unitaries, U rho U^dagger, partial traces,
interweave / replace, measurement with the reference's product-state collapse, ensemble
mixing, and the ProbVal normalise / fan-out ordering rules.  It exists so that the CUDA
path can be checked against something that (a) follows the reference line by line in
*meaning* and (b) is itself pinned to the reference's outputs through the fixtures in
``tests/golden/`` (written by ``tests/golden/make_golden.py``, which imports the real
reference in the build container).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference
arm may import this module, and only as the checker / the timed CPU baseline.  The product
package ``qbot_b200`` never imports it and has no CPU fallback.

Parity status: PINNED.  Every function below is compared with outputs of the reference
itself (``tests/test_oracle_golden.py``), inside the reference's validity domain
(the reference's own swap / multi-control builders are wrong for n >= 5 or
off-slot controls; there the *definitional* unitary is used on both sides and the fixture
says so).

Conventions (reference ``qbot/qgates.py:161-182``): qubit 0 is the most significant bit of
a basis-state index; a k-qubit gate with ``firstTarget = t`` acts on qubits t..t+k-1 and its
own most significant index bit belongs to qubit t.
"""
from __future__ import annotations

import math
from typing import Iterable, List, Sequence, Tuple

import numpy as np

C128 = np.complex128

# ----------------------------------------------------------------------------------------
# shape helpers (reference qbot/helpers.py:9-12, 24-37)
# ----------------------------------------------------------------------------------------

def ilog2(x: int) -> int:
    """floor(log2(x)) with log2(0) == 0, the reference's convention (helpers.py:9-12)."""
    return 0 if x == 0 else int(x).bit_length() - 1


def square_dim(a: np.ndarray) -> int:
    """Side length of a square matrix, 0 for an empty array (helpers.py:24-37)."""
    a = np.asarray(a)
    if a.size == 0:
        return 0
    if a.ndim != 2:
        raise Exception("array must be 2 dimensional")
    if a.shape[0] != a.shape[1]:
        raise Exception("array must be square")
    return a.shape[0]


# ----------------------------------------------------------------------------------------
# full-space unitaries (reference qbot/qgates.py)
# ----------------------------------------------------------------------------------------

def embed_gate(n: int, t: int, g: np.ndarray) -> np.ndarray:
    """I_{2^t} (x) g (x) I_{2^(n-t-k)} -- qgates.py:161-182 (genGateForFullHilbertSpace).

    Raises IndexError when the block does not fit (qgates.py:169-170)."""
    dim = square_dim(g)
    if dim & (dim - 1):
        raise Exception("gate size must be power of 2")
    k = ilog2(dim)
    if t + k - 1 >= n:
        raise IndexError(f"{k} qubit gate does not fit the {n} qubit hilbertspace when started on qubit {t}")
    left = np.eye(1 << t)
    right = np.eye(1 << (n - t - k))
    return np.kron(np.kron(left, g), right)


def controlled_unitary(n: int, controls: Sequence[int], t: int, g: np.ndarray) -> np.ndarray:
    """Definition of the multi-controlled gate that qgates.py:228-275 is meant to build:
    act with ``g`` on qubits t.. when every control qubit is |1>, identity otherwise.

    The reference reaches this matrix by swap / shift conjugations that are only correct in
    part of the parameter space; inside that part the two agree exactly
    (pinned by tests/golden)."""
    full = embed_gate(n, t, g).astype(C128)
    cmask = 0
    for c in controls:
        cmask |= 1 << (n - 1 - c)
    idx = np.arange(1 << n)
    on = (idx & cmask) == cmask
    u = np.eye(1 << n, dtype=C128)
    sel = np.flatnonzero(on)
    u[np.ix_(sel, sel)] = full[np.ix_(sel, sel)]
    return u


def perm_unitary(dim: int, state_map) -> np.ndarray:
    """Permutation matrix sending |i> to |state_map(i)> -- qgates.py:136-141."""
    u = np.zeros((dim, dim), dtype=C128)
    src = np.arange(dim)
    dst = np.array([state_map(int(i)) for i in src])
    u[dst, src] = 1
    return u


def swap_unitary(n: int, a: int, b: int) -> np.ndarray:
    """Definition of the qubit transposition the reference's genSwapGate (qgates.py:77-133)
    is meant to be (it is only correct for n <= 4)."""
    if a == b:
        return np.eye(1 << n, dtype=C128)
    ba, bb = n - 1 - a, n - 1 - b

    def sm(i):
        x = ((i >> ba) ^ (i >> bb)) & 1
        return i ^ ((x << ba) | (x << bb))

    return perm_unitary(1 << n, sm)


def shift_unitary(n: int, up: bool, shifts: int = 1) -> np.ndarray:
    """Cyclic rotation of the index bits -- qgates.py:144-158."""
    dim = 1 << n
    if up:
        sm = lambda s: ((s << shifts) % dim) | ((s << shifts) // dim)
    else:
        sm = lambda s: (s >> shifts) | ((s & ((1 << shifts) - 1)) << (n - shifts))
    return perm_unitary(dim, sm)


def conjugate(u: np.ndarray, rho: np.ndarray) -> np.ndarray:
    """U rho U^dagger -- qgates.py:278-279 (applyGate)."""
    return u @ rho @ u.conj().T


# ----------------------------------------------------------------------------------------
# density helpers (reference qbot/density.py)
# ----------------------------------------------------------------------------------------

def tensor_prod(*parts: np.ndarray) -> np.ndarray:
    """Kronecker chain that skips empty arrays -- density.py:7-24."""
    out = None
    for p in parts:
        p = np.asarray(p)
        if p.size == 0:
            continue
        out = p if out is None else np.kron(out, p)
    return np.array([], dtype=C128) if out is None else out


def ket_to_density(ket: np.ndarray) -> np.ndarray:
    """psi psi^T, *without* conjugation -- density.py:31-32."""
    return np.outer(ket, ket)


def ensemble(probs: Sequence[float], rhos: Sequence[np.ndarray]) -> np.ndarray:
    """sum_b p_b rho_b accumulated left to right -- density.py:49-58."""
    if len(probs) != len(rhos):
        raise Exception("number of state vectors an number of probabilites must equal")
    acc = np.zeros(np.asarray(rhos[0]).shape, dtype=C128)
    for p, r in zip(probs, rhos):
        acc += p * r
    return acc


def _axes_tensor(rho: np.ndarray, n: int) -> np.ndarray:
    return np.asarray(rho, dtype=C128).reshape((2,) * (2 * n))


def ptrace_arbitrary(rho: np.ndarray, n: int, sys_a: Iterable[int]) -> Tuple[np.ndarray, np.ndarray]:
    """(rho_A, rho_B): A = listed qubits (deduplicated, ascending), B = the others
    (ascending); each is the partial trace over the other -- density.py:122-148."""
    square_dim(rho)
    a = sorted(set(int(q) for q in sys_a))
    if a[0] < 0 or a[-1] > n - 1:
        raise IndexError()
    b = [q for q in range(n) if q not in a]
    t = _axes_tensor(rho, n)

    def keep(qs: List[int]) -> np.ndarray:
        if not qs:   # nothing kept: the reference's reshape/trace leaves a 1x1 [[tr rho]]
            return np.array(np.trace(np.asarray(rho)), dtype=C128).reshape(1, 1)
        rows = list(range(n))
        cols = [n + q if q in qs else q for q in range(n)]
        out = [q for q in qs] + [n + q for q in qs]
        r = np.einsum(t, rows + cols, out)
        return r.reshape(1 << len(qs), 1 << len(qs))

    return keep(a), keep(b)


def interweave(rho_a: np.ndarray, rho_b: np.ndarray, a_positions: Iterable[int]) -> np.ndarray:
    """rho_A (x) rho_B with A's qubits placed at the (sorted) positions given and B's
    qubits filling the remaining positions in order -- density.py:150-192."""
    na = ilog2(square_dim(rho_a))
    nb = ilog2(square_dim(rho_b))
    n = na + nb
    pos = sorted(set(int(q) for q in a_positions))
    if pos[0] < 0 or pos[-1] > n - 1:
        raise IndexError(pos, n)
    if len(pos) < na:
        raise ValueError()
    rest = [q for q in range(n) if q not in pos]
    joint = tensor_prod(rho_a, rho_b)
    t = _axes_tensor(joint, n)
    # source qubit order is (A..., B...); final qubit p comes from source axis src[p]
    src = [0] * n
    for i, p in enumerate(pos):
        src[p] = i
    for i, p in enumerate(rest):
        src[p] = na + i
    perm = src + [n + s for s in src]
    return t.transpose(perm).reshape(1 << n, 1 << n)


def replace_arbitrary(rho: np.ndarray, new_rho: np.ndarray, targets: Sequence[int]) -> np.ndarray:
    """Trace the target qubits out and put ``new_rho`` in their place -- density.py:195-227.

    Defined for ASCENDING target lists, which is what the reference's tests use: the i-th qubit
    of ``new_rho`` goes to ``targets[i]``.  For a descending / unsorted list the reference's state
    map (density.py:207-220) is not injective -- it walks the qubits once and only ever matches
    ``qubitsToReplace`` in listed order -- so `genArbitrarySwap` returns a non-unitary 0/1 matrix
    and the result is not a density matrix (finding F13; found by
    scripts/fuzz_dsl.py).  There the positions are taken as a set (sorted), like
    `interweaveDensities` does: parity unpinned for unsorted lists."""
    n = ilog2(square_dim(rho))
    k = ilog2(square_dim(new_rho))
    if len(targets) != k:
        raise ValueError(f'number of target qubits {len(targets)} does not equal number of provided qubits {k}')
    _, rest = ptrace_arbitrary(rho, n, targets)      # 1x1 [[tr rho]] when n == k
    return interweave(new_rho, rest, targets)


# ----------------------------------------------------------------------------------------
# measurement (reference qbot/measurement.py:88-165 and 18-29)
# ----------------------------------------------------------------------------------------

PROB_ROUNDING = 15   # probVal.py:8
SMALL_VAL = 1e-5     # probVal.py:7


def basis_projector(num_factors: int, index: int, basis_density: Sequence[np.ndarray],
                    symbols: Sequence[str] = None):
    """index-th tensor permutation of the basis projectors, most significant digit first
    -- measurement.py:88-101 (permuteBasis)."""
    base = len(basis_density)
    digits = []
    rem = index
    for _ in range(num_factors):
        digits.append(rem % base)
        rem //= base
    digits.reverse()
    proj = tensor_prod(*[basis_density[d] for d in digits])
    sym = ''.join(symbols[d] for d in digits) if symbols is not None else ''
    return proj, sym


def basis_weights(state: np.ndarray, n: int, targets: Sequence[int], kets: Sequence[np.ndarray]) -> np.ndarray:
    """|tr(rho_A P_i)| for every outcome i of measurement.py:147-155 WITHOUT forming a projector:
    P_i = (x)_f outer(ket_{d_f}, ket_{d_f}) (no conjugation, basis.py:24-26 / density.py:31-32), so
    tr(rho_A P_i) = the i-th diagonal entry of (W x..x W) rho_A (W x..x W)^T with the basis kets as the
    rows of W.  `state` is a density matrix (2-D) or a ket (1-D: weights sum |W psi|^2, equal to the
    density form for real bases).  `targets` ascending; groups of b = log2(len(kets[0])) qubits, first
    group = most significant digit.  Checked against the literal loop (`measure`) in the CPU tests;
    this form stays feasible at 9-13 measured qubits where the loop's 2^k x 2^k matmuls do not."""
    W = np.stack([np.asarray(k, dtype=C128).reshape(-1) for k in kets])
    b = ilog2(W.shape[1])
    m = len(targets)
    assert m % b == 0 and list(targets) == sorted(targets)
    Wt = W.reshape((W.shape[0],) + (2,) * b)

    def rotate(t, axes):
        """contract the basis kets with the listed axes of t; outcome digit bits go back to the same axes"""
        r = np.tensordot(Wt, t, axes=(list(range(1, b + 1)), list(axes)))
        r = r.reshape((2,) * b + r.shape[1:])
        return np.moveaxis(r, list(range(b)), list(axes))

    a = np.asarray(state, dtype=C128)
    if a.ndim == 1:
        t = a.reshape((2,) * n)
        for f in range(m // b):
            t = rotate(t, targets[f * b:(f + 1) * b])
        p = np.abs(t) ** 2
        others = tuple(q for q in range(n) if q not in targets)
        return (p.sum(axis=others) if others else p).reshape(-1)
    rho = a if m == n else ptrace_arbitrary(a, n, list(targets))[0]
    t = rho.reshape((2,) * (2 * m))
    for f in range(m // b):
        rows = list(range(f * b, (f + 1) * b))
        t = rotate(t, rows)
        t = rotate(t, [m + r for r in rows])
    return np.abs(np.diag(t.reshape(1 << m, 1 << m)))


def measure(rho: np.ndarray, basis_density: Sequence[np.ndarray], targets=None,
            return_state: bool = True, symbols: Sequence[str] = None) -> dict:
    """The reference's measurement -- measurement.py:107-165 plus the MeasurementResult
    constructor 18-29.  Returns a dict with the MeasurementResult slots.

    Collapse is the reference's *product state* interweave(sum_i p_i P_i, Tr_A rho)
    (SURVEY.md F7), not sum_i P_i rho P_i."""
    n = ilog2(square_dim(rho))
    if targets is None:
        tlist = None
        num_targets = n
    else:
        tlist = list(targets) if isinstance(targets, set) else list(set(targets))
        for q in tlist:
            if q < 0 or q > n - 1:
                raise IndexError(f"measurement target {q} outside of valid range [0, {n - 1}]")
        num_targets = len(tlist)
    bq = ilog2(square_dim(basis_density[0]))
    if num_targets == 0:
        raise ValueError("measurement must have targets")
    if num_targets % bq != 0:
        raise ValueError(f"number of qubits to measure {num_targets} must be divisable by the number of qubits in the basis states {bq}")
    if tlist is None or len(tlist) == n:
        sys_a, sys_b = np.asarray(rho), np.array([], dtype=C128)
    else:
        sys_a, sys_b = ptrace_arbitrary(rho, n, tlist)
    factors = num_targets // bq
    probs, projs, syms = [], [], []
    total = 0
    for i in range(len(basis_density) ** factors):
        proj, sym = basis_projector(factors, i, basis_density, symbols)
        probs.append(abs(np.trace(np.matmul(sys_a, proj))))
        projs.append(proj)
        syms.append(sym)
        total += probs[-1]
    probs = [p / total for p in probs]
    new_state = None
    if return_state:
        measured = ensemble(probs, projs)
        new_state = measured if tlist is None else interweave(measured, sys_b, tlist)
    # MeasurementResult.__init__ renormalises and rounds (measurement.py:22-25)
    s = sum(probs)
    probs = [round(p / s, PROB_ROUNDING) for p in probs]
    return dict(unMeasuredDensity=sys_a, probs=probs, basisDensity=projs, basisSymbols=syms,
                newState=new_state)


def run_config3(ops, n: int):
    """BASELINE config 3 in the reference's semantics, op by op: `ops` is a list of
    records with .kind in {'gate', 'meas', 'pgate', 'disc'}, .gate (target / controls / matrix()),
    .qubits and .name (qbot_b200.circuits.c3_ops).  gate = operators.py:255-329 (U rho U^dagger),
    meas = measurement.py:107-165 (computational basis, product-state collapse F7), pgate = a Hadamard on
    a ProbVal([.5,.5]) target (operators.py:308-316: per-branch applyGate + ensemble), disc =
    operators.py:169-175 (keeps the qubits that are not listed).  Returns (final rho, {name: probs})."""
    rho = np.zeros((1 << n, 1 << n), dtype=C128)
    rho[0, 0] = 1
    comp = [np.diag([1, 0]).astype(C128), np.diag([0, 1]).astype(C128)]
    had = np.array([[1, 1], [1, -1]], dtype=C128) * 2 ** (-1 / 2)
    probs = {}
    for op in ops:
        if op.kind == 'gate':
            rho = dm_apply(rho, n, op.gate.target, op.gate.matrix(), op.gate.controls)
        elif op.kind == 'meas':
            r = measure(rho, comp, list(op.qubits), True)
            probs[op.name] = r['probs']
            rho = r['newState']
        elif op.kind == 'pgate':
            rho = ensemble([.5, .5], [dm_apply(rho, n, t, had, []) for t in op.qubits])
        else:
            rho = ptrace_arbitrary(rho, n, list(op.qubits))[1]
    return rho, probs


# ----------------------------------------------------------------------------------------
# ProbVal rules (reference qbot/probVal.py:22-51, 347-390)
# ----------------------------------------------------------------------------------------

def _vals_close(a, b) -> bool:
    if isinstance(a, float):
        return abs(a - b) < SMALL_VAL
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        return bool((a == b).all())
    return a == b


def probval_normalize(probs: Sequence[float], values: Sequence) -> Tuple[List[float], list]:
    """probVal.py:22-51: drop p < 1e-5, drop (not merge) later duplicates, renormalise,
    round to 15 decimals."""
    p, v = list(probs), list(values)
    i = 0
    while i < len(p):
        if p[i] < SMALL_VAL:
            del p[i], v[i]
            continue
        j = i + 1
        while j < len(p):
            if _vals_close(v[i], v[j]):
                del p[j], v[j]
            else:
                j += 1
        i += 1
    s = sum(p)
    p = [round(x / s, PROB_ROUNDING) for x in p]
    return p, v


def fan_out(arg_lens: Sequence[int]) -> List[Tuple[int, ...]]:
    """Branch enumeration order of funcWrapper (probVal.py:347-390): the Cartesian product
    of all ProbVal arguments with the FIRST ProbVal argument varying fastest (F10).
    ``arg_lens[i]`` is the number of branches of the i-th ProbVal argument."""
    total = 1
    for l in arg_lens:
        total *= l
    out = []
    for it in range(total):
        rem = it
        pick = []
        for l in arg_lens:
            pick.append(rem % l)
            rem //= l
        out.append(tuple(pick))
    return out


# ----------------------------------------------------------------------------------------
# ket-level restatement (not a reference code path -- the reference has no ket path,
# SURVEY.md F1; parity target for kets is rho_ref == outer(psi, conj psi))
# ----------------------------------------------------------------------------------------

def ket_apply(psi: np.ndarray, n: int, t: int, g: np.ndarray, controls: Sequence[int] = ()) -> np.ndarray:
    """controlled_unitary(n, controls, t, g) @ psi without building the 2^n x 2^n matrix."""
    g = np.asarray(g, dtype=C128)
    k = ilog2(g.shape[0])
    if t + k - 1 >= n or t < 0:
        raise IndexError("gate does not fit")
    out = np.array(psi, dtype=C128).reshape((2,) * n)
    sl = [slice(None)] * n
    for c in controls:
        sl[c] = 1
    sub = out[tuple(sl)]
    # axes of `sub` corresponding to the target qubits: controls removed before them shift
    ctl = sorted(controls)
    axes = [q - sum(1 for c in ctl if c < q) for q in range(t, t + k)]
    gt = g.reshape((2,) * (2 * k))
    res = np.tensordot(gt, sub, axes=(list(range(k, 2 * k)), axes))
    res = np.moveaxis(res, list(range(k)), axes)
    out[tuple(sl)] = res
    return out.reshape(-1)


def ket_apply_inplace(psi: np.ndarray, n: int, t: int, g: np.ndarray, controls: Sequence[int] = ()) -> np.ndarray:
    """ket_apply for registers of 2^26 and more amplitudes: the same two products and one sum per output
    amplitude, on strided views of `psi` itself (a complex128 array that is overwritten) -- no full-size
    temporaries.  1-qubit gates only; larger blocks go through ket_apply.  Checked against ket_apply in the
    CPU tests."""
    g = np.asarray(g, dtype=C128)
    if g.shape != (2, 2):
        psi[...] = ket_apply(psi, n, t, g, controls)
        return psi
    if t < 0 or t >= n:
        raise IndexError("gate does not fit")
    T = psi.reshape((2,) * n)
    sl = [slice(None)] * n
    for c in controls:
        sl[c] = 1
    s0, s1 = list(sl), list(sl)
    s0[t], s1[t] = 0, 1
    a0, a1 = T[tuple(s0)], T[tuple(s1)]
    if g[0, 1] == 0 and g[1, 0] == 0:
        if g[0, 0] != 1:
            a0 *= g[0, 0]
        if g[1, 1] != 1:
            a1 *= g[1, 1]
        return psi
    n0 = g[0, 0] * a0
    n0 += g[0, 1] * a1
    a1 *= g[1, 1]
    a1 += g[1, 0] * a0
    a0[...] = n0
    return psi


def ket_swap(psi: np.ndarray, n: int, a: int, b: int) -> np.ndarray:
    return np.array(psi, dtype=C128).reshape((2,) * n).swapaxes(a, b).reshape(-1).copy()


def ket_probs(psi: np.ndarray, n: int, targets: Sequence[int]) -> np.ndarray:
    """Computational-basis outcome probabilities of the listed qubits (first listed qubit
    = most significant outcome bit)."""
    p = (np.abs(np.asarray(psi)) ** 2).reshape((2,) * n)
    others = tuple(q for q in range(n) if q not in targets)
    p = p.sum(axis=others) if others else p
    # axes now ordered ascending by qubit; reorder to the listed order
    asc = sorted(targets)
    p = np.transpose(p, [asc.index(q) for q in targets])
    return p.reshape(-1)


def ket_density(psi: np.ndarray) -> np.ndarray:
    """psi psi^dagger -- the physical density matrix of a ket."""
    return np.outer(psi, np.conj(psi))


def dm_apply(rho: np.ndarray, n: int, t: int, g: np.ndarray, controls: Sequence[int] = ()) -> np.ndarray:
    """U rho U^dagger with U = controlled_unitary(n, controls, t, g), computed on rho as a
    2n-index tensor (for sizes where the 2^n x 2^n unitary is too slow to build)."""
    dim = 1 << n
    flat = np.asarray(rho, dtype=C128).reshape(-1)
    # rows: act with g on qubit t of the first n axes
    flat = ket_apply(flat, 2 * n, t, g, controls)
    # columns: act with conj(g) on the second n axes
    flat = ket_apply(flat, 2 * n, n + t, np.conj(g), [n + c for c in controls])
    return flat.reshape(dim, dim)


# ----------------------------------------------------------------------------------------
# the reference's algorithm, for the timed CPU baseline (bench.py --impl reference)
# ----------------------------------------------------------------------------------------

def reference_style_gate(rho: np.ndarray, n: int, t: int, g: np.ndarray, controls: Sequence[int] = ()) -> np.ndarray:
    """One `gate` op the way the reference computes it: materialise the 2^n x 2^n unitary
    (kron padding, qgates.py:161-182; control structure per 228-275) and do two dense
    matmuls (qgates.py:278-279).  Cost model: 16*8^n flop + O(4^n) bytes per gate."""
    u = embed_gate(n, t, g) if len(controls) == 0 else controlled_unitary(n, controls, t, g)
    return conjugate(u, rho)
