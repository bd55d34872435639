"""numpy restatement of lduMatrix::operator+= / -= / *= (LDU/lduMatrix/lduMatrixOperations.C:235-465) on
optional diag / upper / lower arrays -- TEST INFRASTRUCTURE (checker of b200ldu_ldu_add_assign / _scale).
Pinned: tests/test_lduops.py runs the reference's own operators (oracle/_ref/libref_lduops.so) beside it for
every kind combination."""
import numpy as np


def _kinds(m):
    d, u, l = (m.get(k) is not None for k in ("diag", "upper", "lower"))
    return d and u and not l, d and u and l, d and not u and not l   # symmetric, asymmetric, diagonal (lduMatrix.H:626-639)


def add_assign(A, B, nCells, sub=False):
    A = {k: (None if v is None else np.array(v, dtype=np.float64)) for k, v in A.items()}
    op = (lambda x, y: x - y) if sub else (lambda x, y: x + y)
    if B.get("diag") is not None:
        A["diag"] = op(A["diag"] if A.get("diag") is not None else np.zeros(nCells), B["diag"])
    symA, asymA, diagA = _kinds(A)
    symB, asymB, _ = _kinds(B)
    if symA and symB:
        A["upper"] = op(A["upper"], B["upper"])
    elif symA and asymB:
        A["lower"] = A["upper"].copy()
        A["upper"] = op(A["upper"], B["upper"])
        A["lower"] = op(A["lower"], B["lower"])
    elif asymA and symB:
        A["lower"] = op(A["lower"], B["upper"])
        A["upper"] = op(A["upper"], B["upper"])
    elif asymA and asymB:
        A["lower"] = op(A["lower"], B["lower"])
        A["upper"] = op(A["upper"], B["upper"])
    elif diagA:
        if B.get("upper") is not None:
            A["upper"] = -B["upper"] if sub else B["upper"].copy()
        if B.get("lower") is not None:
            A["lower"] = -B["lower"] if sub else B["lower"].copy()
    return {k: v for k, v in A.items() if v is not None}


def scale(A, s, lower, upper):
    out = {}
    field = not np.isscalar(s)
    if A.get("diag") is not None:
        out["diag"] = A["diag"] * s
    if A.get("upper") is not None:
        out["upper"] = A["upper"] * (s[np.asarray(lower)] if field else s)
    if A.get("lower") is not None:
        out["lower"] = A["lower"] * (s[np.asarray(upper)] if field else s)
    return out
