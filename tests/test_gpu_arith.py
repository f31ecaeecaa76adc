"""The exact fast FP64 division / sqrt (csrc/rt_arith.cuh) must reproduce the native operators
bit for bit wherever it declares itself valid, and must declare itself valid almost always."""
import ctypes as C

import numpy as np
import pytest

from ray_tracer_challenge_rs_b200 import abi

pytestmark = pytest.mark.gpu


def run(a, b):
    lib = abi.load_library()
    a = np.ascontiguousarray(a, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    out = [C.c_uint64() for _ in range(4)]
    st = lib.rtgpu_selftest_arith(0, a.ctypes.data_as(C.POINTER(C.c_double)), b.ctypes.data_as(C.POINTER(C.c_double)), a.size,
                                  *[C.byref(o) for o in out])
    abi.check(lib, st)
    return [o.value for o in out]


def test_random_operands_bit_exact():
    rng = np.random.default_rng(0xB200)
    n = 1 << 24
    # magnitudes spread over the whole range a render produces (and far beyond), both signs
    a = rng.standard_normal(n) * np.exp2(rng.integers(-60, 60, n))
    b = rng.standard_normal(n) * np.exp2(rng.integers(-60, 60, n))
    a[: n // 2] = np.abs(a[: n // 2])  # sqrt operands
    div_bad, sqrt_bad, div_fb, sqrt_fb = run(a, b)
    assert div_bad == 0 and sqrt_bad == 0
    assert div_fb < n * 1e-6, div_fb
    assert sqrt_fb <= n // 2 + n // 1000, sqrt_fb  # the negative half takes the native path (NaN)


def test_mantissa_corner_cases_bit_exact():
    """Operands with all-ones / all-zeros mantissas and neighbours of powers of two: where a
    Newton-Raphson division is most likely to misround."""
    mant = np.array([0, 1, 2, 3, (1 << 52) - 1, (1 << 52) - 2, (1 << 51), (1 << 51) - 1, (1 << 51) + 1, (1 << 26), (1 << 26) - 1,
                     0x5555555555555, 0xAAAAAAAAAAAAA, 0xFFFFF00000000, 0x00000FFFFFFFF], dtype=np.uint64)
    exps = np.array([1023 - 40, 1023 - 1, 1023, 1023 + 1, 1023 + 40], dtype=np.uint64)
    vals = ((exps[:, None] << np.uint64(52)) | mant[None, :]).reshape(-1).view(np.float64)
    vals = np.concatenate([vals, -vals])
    a, b = np.meshgrid(vals, vals)
    div_bad, sqrt_bad, div_fb, sqrt_fb = run(a.ravel(), b.ravel())
    assert div_bad == 0 and sqrt_bad == 0 and div_fb == 0


def test_special_operands_fall_back():
    specials = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 5e-324, 2.2e-308, 1.7e308, 1e-300, 1e300, 1.0, -3.5])
    a, b = np.meshgrid(specials, specials)
    div_bad, sqrt_bad, div_fb, sqrt_fb = run(a.ravel(), b.ravel())
    assert div_bad == 0 and sqrt_bad == 0
    assert div_fb > 0 and sqrt_fb > 0
