"""Host-side math primitives, restated so that every matrix the flattener emits is bit-identical
to what the reference holds.

Mirrors (reference paths under ``ray-tracer/src/primitives/``):
  * ``Matrix<4>`` with cofactor inverse             -> matrix.rs:8-258, 317-330
  * transformation builders + ``view_transform``    -> transformations.rs:5-87
  * ``Vector::{dot,cross,normalized,magnitude}``     -> vector.rs:84-103
  * ``EPSILON`` / ``CoarseEq``                       -> consts.rs:2, utils.rs:16-24

Only the once-per-scene host work lives here (builders, inverse, view transform).  The per-ray
arithmetic is the CUDA kernels' job; nothing in this module renders.

Plain Python floats are IEEE binary64 and CPython never contracts ``a*b+c``, so expression order
here *is* the reference's order.  ``mul_add`` sites use a correctly rounded :func:`fma`.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import math
from fractions import Fraction
from typing import Iterable, List, Sequence, Tuple

EPSILON = 0.00000008  # consts.rs:2
F64_MAX = 1.7976931348623157e308  # consts.rs:6
F64_MIN = -F64_MAX  # consts.rs:4

Vec3 = Tuple[float, float, float]
Matrix = List[List[float]]  # row-major, matrix.rs:8


def _load_libm_fma():
    try:
        libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
        f = libm.fma
        f.restype = ctypes.c_double
        f.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_double]
        # sanity: a case where fused != unfused
        if f(1.0 + 2.0**-30, 1.0 - 2.0**-30, -1.0) != -(2.0**-60):
            return None
        return f
    except Exception:  # pragma: no cover - libm is always there on Linux
        return None


_libm_fma = _load_libm_fma()


def fma(a: float, b: float, c: float) -> float:
    """IEEE fusedMultiplyAdd (Rust ``f64::mul_add``).  libm when available, else exact rationals."""
    if _libm_fma is not None:
        return _libm_fma(a, b, c)
    if not (math.isfinite(a) and math.isfinite(b) and math.isfinite(c)):
        return a * b + c
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def coarse_eq(a: float, b: float) -> bool:
    """utils.rs:16-24"""
    if a == b:
        return True
    return abs(a - b) < EPSILON


# --------------------------------------------------------------------------------------------
# Vector (vector.rs)


def dot(a: Sequence[float], b: Sequence[float]) -> float:
    """vector.rs:93-95"""
    return fma(a[2], b[2], fma(a[0], b[0], a[1] * b[1]))


def cross(a: Sequence[float], b: Sequence[float]) -> Vec3:
    """vector.rs:97-103"""
    return (
        fma(a[1], b[2], -a[2] * b[1]),
        fma(a[2], b[0], -a[0] * b[2]),
        fma(a[0], b[1], -a[1] * b[0]),
    )


def magnitude(a: Sequence[float]) -> float:
    """vector.rs:84-86"""
    return math.sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2])


def normalized(a: Sequence[float]) -> Vec3:
    """vector.rs:88-91 (divides, does not multiply by a reciprocal)"""
    m = magnitude(a)
    return (a[0] / m, a[1] / m, a[2] / m)


def sub(a: Sequence[float], b: Sequence[float]) -> Vec3:
    return (a[0] - b[0], a[1] - b[1], a[2] - b[2])


# --------------------------------------------------------------------------------------------
# Matrix<N> (matrix.rs)


def identity(n: int = 4) -> Matrix:
    return [[1.0 if r == c else 0.0 for c in range(n)] for r in range(n)]


def matrix(rows: Iterable[Iterable[float]]) -> Matrix:
    return [[float(v) for v in row] for row in rows]


def transpose(m: Matrix) -> Matrix:
    """matrix.rs:30-43"""
    n = len(m)
    return [[m[c][r] for c in range(n)] for r in range(n)]


def is_identity(m: Matrix) -> bool:
    """matrix.rs:45-51"""
    n = len(m)
    return all(coarse_eq(m[r][c], 1.0 if r == c else 0.0) for r in range(n) for c in range(n))


def mat_mul(a: Matrix, b: Matrix) -> Matrix:
    """matrix.rs:317-330: fold from 0.0 over k"""
    n = len(a)
    out = [[0.0] * n for _ in range(n)]
    for r in range(n):
        for c in range(n):
            acc = 0.0
            for k in range(n):
                acc = acc + (a[r][k] * b[k][c])
            out[r][c] = acc
    return out


def mat_point(m: Matrix, p: Sequence[float]) -> Vec3:
    """matrix.rs:332-346: fold from 0.0 over [x, y, z, 1.0], rows 0..2"""
    vals = (p[0], p[1], p[2], 1.0)
    out = []
    for r in range(3):
        acc = 0.0
        for c in range(4):
            acc = acc + (m[r][c] * vals[c])
        out.append(acc)
    return (out[0], out[1], out[2])


def mat_vector(m: Matrix, v: Sequence[float]) -> Vec3:
    """matrix.rs:348-362: fold from 0.0 over [x, y, z, 0.0]"""
    vals = (v[0], v[1], v[2], 0.0)
    out = []
    for r in range(3):
        acc = 0.0
        for c in range(4):
            acc = acc + (m[r][c] * vals[c])
        out.append(acc)
    return (out[0], out[1], out[2])


def _submatrix(m: Matrix, excluded_row: int, excluded_column: int) -> Matrix:
    """matrix.rs:120-144 (3x3), 190-218 (4x4)"""
    n = len(m)
    if n == 4 and is_identity(m):  # matrix.rs:191-193 (quirk kept)
        return identity(3)
    return [
        [m[r][c] for c in range(n) if c != excluded_column] for r in range(n) if r != excluded_row
    ]


def determinant(m: Matrix) -> float:
    n = len(m)
    if n == 2:  # matrix.rs:83-85
        return (m[0][0] * m[1][1]) - (m[0][1] * m[1][0])
    acc = 0.0  # matrix.rs:151-155, 223-227
    for i in range(n):
        acc = acc + (m[0][i] * cofactor(m, 0, i))
    return acc


def minor(m: Matrix, row: int, column: int) -> float:
    n = len(m)
    if n == 2:  # matrix.rs:65-80, 87-89: the remaining single element
        return m[1 - row][1 - column]
    return determinant(_submatrix(m, row, column))


def cofactor(m: Matrix, row: int, column: int) -> float:
    """matrix.rs:91-98, 161-168, 233-240"""
    mn = minor(m, row, column)
    return mn if (row + column) % 2 == 0 else -mn


def inverse(m: Matrix) -> Matrix:
    """matrix.rs:104-117, 174-186, 246-258: exact IDENTITY when identity within EPSILON, else
    cofactor(column, row) / determinant."""
    n = len(m)
    if is_identity(m):
        return identity(n)
    det = determinant(m)
    out = [[0.0] * n for _ in range(n)]
    for r in range(n):
        for c in range(n):
            cf = cofactor(m, c, r)
            out[r][c] = cf / det if det != 0.0 else _div(cf, det)
    return out


def _div(a: float, b: float) -> float:
    """IEEE division including division by zero (Python raises instead)."""
    if b != 0.0:
        return a / b
    if a != a or a == 0.0:
        return math.nan
    neg = (math.copysign(1.0, a) < 0) != (math.copysign(1.0, b) < 0)
    return -math.inf if neg else math.inf


# --------------------------------------------------------------------------------------------
# transformations.rs


def translation(x: float, y: float, z: float) -> Matrix:
    """transformations.rs:5-11"""
    m = identity()
    m[0][3] = float(x)
    m[1][3] = float(y)
    m[2][3] = float(z)
    return m


def scaling(x: float, y: float, z: float) -> Matrix:
    """transformations.rs:13-19"""
    m = identity()
    m[0][0] = float(x)
    m[1][1] = float(y)
    m[2][2] = float(z)
    return m


def rotation_x(theta: float) -> Matrix:
    """transformations.rs:21-31"""
    theta = float(theta)
    m = identity()
    cos, sin = math.cos(theta), math.sin(theta)
    m[1][1] = cos
    m[1][2] = -sin
    m[2][1] = sin
    m[2][2] = cos
    return m


def rotation_y(theta: float) -> Matrix:
    """transformations.rs:33-43"""
    theta = float(theta)
    m = identity()
    cos, sin = math.cos(theta), math.sin(theta)
    m[0][0] = cos
    m[0][2] = sin
    m[2][0] = -sin
    m[2][2] = cos
    return m


def rotation_z(theta: float) -> Matrix:
    """transformations.rs:45-55"""
    theta = float(theta)
    m = identity()
    cos, sin = math.cos(theta), math.sin(theta)
    m[0][0] = cos
    m[0][1] = -sin
    m[1][0] = sin
    m[1][1] = cos
    return m


def shearing(xy: float, xz: float, yx: float, yz: float, zx: float, zy: float) -> Matrix:
    """transformations.rs:57-73"""
    m = identity()
    m[0][1] = float(xy)
    m[0][2] = float(xz)
    m[1][0] = float(yx)
    m[1][2] = float(yz)
    m[2][0] = float(zx)
    m[2][1] = float(zy)
    return m


def view_transform(from_: Sequence[float], to: Sequence[float], up: Sequence[float]) -> Matrix:
    """transformations.rs:75-87"""
    forward = normalized(sub(to, from_))
    up_normalized = normalized(up)
    left_vector = cross(forward, up_normalized)
    true_up = cross(left_vector, forward)
    orientation = [
        [left_vector[0], left_vector[1], left_vector[2], 0.0],
        [true_up[0], true_up[1], true_up[2], 0.0],
        [-forward[0], -forward[1], -forward[2], 0.0],
        [0.0, 0.0, 0.0, 1.0],
    ]
    return mat_mul(orientation, translation(-from_[0], -from_[1], -from_[2]))
