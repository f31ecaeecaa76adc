"""The single-precision bounding-sphere pre-test of the f64 kernels (rt_kernel.cuh trace_unified, RT_CULL_F32) must be
SAFE: it may only discard a shape when, in exact arithmetic, the ray's line misses the (double precision) cull sphere or
the sphere lies wholly behind the origin.  This mirrors the kernel's arithmetic in numpy float32 (fma emulated through
float64: the product of two floats is exact there) and the packer's record (rtgpu.cu pack_scene: centre rounded to
nearest, radius^2 padded by 2^-9 and rounded up), and checks the claim against rational arithmetic on rays aimed at the
rim of the sphere, where the margins matter.
"""
from fractions import Fraction

import numpy as np

f32 = np.float32


def fma32(a, b, c):
    return f32(np.float64(a) * np.float64(b) + np.float64(c))


def record32(c, r2):
    padded = r2 * (1.0 + 2.0 ** -9)
    r = f32(padded)
    if float(r) < padded:
        r = np.nextafter(r, f32(np.inf))
    return [f32(v) for v in c] + [r]


def culls32(o, d, rec, coord_max, container=False):
    """rt_kernel.cuh trace_unified, RT_CULL_F32 branch, one shape."""
    ox, oy, oz = (f32(v) for v in o)
    dx, dy, dz = (f32(v) for v in d)
    dd = fma32(dz, dz, fma32(dy, dy, f32(dx * dx)))
    m = max(abs(ox), abs(oy), abs(oz), f32(coord_max))
    pad = f32(f32(m * m) * f32(2.0 ** -29))
    if not (dd > f32(1e-30) and dd < f32(1e30)) or not (m < f32(1e18)):
        pad = f32(np.inf)
    behind_below = f32(-np.inf) if container else f32(0)
    shrink = f32(1.0) - f32(2.0 ** -18)
    cx, cy, cz, r2 = rec
    ocx, ocy, ocz = f32(cx - ox), f32(cy - oy), f32(cz - oz)
    bq = fma32(ocz, dz, fma32(ocy, dy, f32(ocx * dx)))
    c2 = fma32(ocz, ocz, fma32(ocy, ocy, f32(ocx * ocx)))
    with np.errstate(invalid="ignore", over="ignore"):
        ex = fma32(c2, shrink, -f32(r2 + pad))
        outside, behind, misses = ex > 0, bq < behind_below, f32(ex * dd) > f32(bq * bq)
    return bool(outside and (behind or misses))


def truly_irrelevant(o, d, c, r2, container=False):
    """Exact: the line misses the sphere, or (not a container walk) origin outside and centre behind."""
    F = Fraction
    oc = [F(ci) - F(oi) for ci, oi in zip(c, o)]
    dq = [F(v) for v in d]
    bq = sum(a * b for a, b in zip(oc, dq))
    c2 = sum(a * a for a in oc)
    dd = sum(a * a for a in dq)
    misses = c2 * dd - bq * bq > F(r2) * dd
    behind = (not container) and c2 > F(r2) and bq < 0
    return misses or behind


def rim_rays(rng, n, scale):
    """Rays whose line passes within a relative 1e-9 .. 1e-3 of the sphere's rim, from near and far, plus random ones."""
    for _ in range(n):
        c = rng.uniform(-scale, scale, 3)
        r = float(10 ** rng.uniform(-2, 1.5))
        o = c + rng.normal(size=3) * r * float(10 ** rng.uniform(0.01, 3))
        toward = c - o
        dist = np.linalg.norm(toward)
        u = np.cross(toward, rng.normal(size=3))
        u /= np.linalg.norm(u)
        rim = c + u * r * (1.0 + rng.choice([-1, 1]) * float(10 ** rng.uniform(-9, -3)))
        d = rim - o
        if rng.random() < 0.7:
            d /= np.linalg.norm(d)  # the reference's rays are unit, refracted ones nearly
        if rng.random() < 0.3:
            d = -d  # sphere behind the origin
        if rng.random() < 0.15 and dist > 0:
            o = c + rng.normal(size=3) * r * 0.5  # origin inside
        yield o, d, c, r * r


def test_single_precision_pretest_never_discards_a_relevant_shape():
    rng = np.random.default_rng(20261018)
    culled = total = 0
    for scale in (1.0, 30.0, 1.0e4):
        for o, d, c, r2 in rim_rays(rng, 700, scale):
            rec = record32(c, r2)
            coord_max = np.nextafter(f32(np.max(np.abs(c))), f32(np.inf))
            for container in (False, True):
                total += 1
                if culls32(o, d, rec, coord_max, container):
                    culled += 1
                    assert truly_irrelevant(o, d, c, r2, container), (o, d, c, r2, container)
    assert culled > total // 10  # not vacuous: even among rim-grazing rays the pre-test discards a good part


def test_single_precision_pretest_degenerate_inputs_keep_the_shape():
    rec = record32([1.0, 2.0, 3.0], 4.0)
    far = [50.0, 50.0, 50.0]
    assert culls32(far, [1.0, 0.0, 0.0], rec, 3.0)  # sanity: a plain miss is culled
    for d in ([0.0, 0.0, 0.0], [1e-20, 0.0, 0.0], [1e20, 0.0, 0.0], [np.nan, 0.0, 1.0], [np.inf, 0.0, 0.0]):
        assert not culls32(far, d, rec, 3.0), d
    for o in ([np.nan, 0.0, 0.0], [1e19, 0.0, 0.0], [np.inf, 0.0, 0.0]):
        assert not culls32(o, [1.0, 0.0, 0.0], rec, 3.0), o
    assert not culls32(far, [1.0, 0.0, 0.0], record32([0.0, 0.0, 0.0], np.inf), 0.0)  # a plane's record
    assert not culls32(far, [1.0, 0.0, 0.0], rec, np.inf)  # an unrepresentable centre somewhere in the scene
