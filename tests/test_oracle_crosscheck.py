"""CPU: an independent restatement of the coverage rules, checked against the C oracle on random triangles.

The reference holds no known answers for rasterisation, so the oracle is the only definition of "correct"
the GPU kernels are held to.  This file derives coverage a second time, straight from the Vulkan 1.1 text
("Basic Polygon Rasterization": fixed-point vertices with subPixelPrecisionBits = 4, sample at the pixel centre, a sample
on an edge belongs to the triangle only if that edge is a top or a left edge) and in a different way from oracle/raster.c:

* exact integers in 1/16-pixel units for everything (no floating point after the snap);
* "top" and "left" are decided GEOMETRICALLY from the position of the opposite vertex (a top edge is horizontal with the
  triangle below it; a left edge is non-horizontal with the triangle to its right), not from the signs of the
  orientation-normalised edge coefficients the oracle uses;
* inside/outside by comparing the sample's side of each edge with the opposite vertex's side.

Triangles are drawn on a coarse vertex grid so that many samples fall exactly on edges and vertices.
"""
import numpy as np
import pytest

from tests.test_oracle_raster import covered, tri
from tests.util import Scene

SIZE = 32  # a power of two: pixel -> NDC -> window is exact in fp32 for the 1/32-pixel positions used below


def snap(v: float) -> int:
    """Window coordinate -> 1/16 pixel units, round to nearest even (the rule of the fixed-point snap, E.6)."""
    return int(np.rint(np.float32(v) * np.float32(16.0)))  # np.rint rounds half to even, as cvt.rni / lrintf do


def side(ax, ay, bx, by, px, py):
    """Twice the signed area of (A, B, P): which side of the line A->B the point is on (exact integers)."""
    return (bx - ax) * (py - ay) - (by - ay) * (px - ax)


def owns_its_boundary(ax, ay, bx, by, cx, cy) -> bool:
    """Is edge AB of triangle ABC (C the opposite vertex, y pointing down) a top or a left edge?"""
    if ay == by:
        return cy > ay  # horizontal: top edge if the triangle is below it
    # non-horizontal: x of the edge's line at the height of C, compared with C (cross-multiplied: exact)
    # x_edge = ax + (cy - ay) * (bx - ax) / (by - ay);  left edge  <=>  cx > x_edge
    num = (cx - ax) * (by - ay) - (cy - ay) * (bx - ax)
    return num > 0 if by > ay else num < 0


def reference_coverage(points_px, size=SIZE) -> np.ndarray:
    pts = [(snap(x), snap(y)) for x, y in points_px]
    (x0, y0), (x1, y1), (x2, y2) = pts
    out = np.zeros((size, size), bool)
    if side(x0, y0, x1, y1, x2, y2) == 0:
        return out  # zero area after snapping: no fragments
    edges = [(pts[0], pts[1], pts[2]), (pts[1], pts[2], pts[0]), (pts[2], pts[0], pts[1])]
    for py in range(size):
        for px in range(size):
            sx, sy = 16 * px + 8, 16 * py + 8
            inside = True
            for (a, b, c) in edges:
                s_p = side(*a, *b, sx, sy)
                s_c = side(*a, *b, *c)
                if s_p == 0:
                    inside = inside and owns_its_boundary(*a, *b, *c)
                else:
                    inside = inside and ((s_p > 0) == (s_c > 0))
            out[py, px] = inside
    return out


def _random_triangles(n, seed):
    rng = np.random.default_rng(seed)
    for k in range(n):
        if k % 3 == 0:   # vertices on half-pixel positions: edges and vertices through pixel centres
            pts = rng.integers(-2, 2 * SIZE + 4, (3, 2)) / 2.0
        elif k % 3 == 1:  # vertices on the 1/16 grid
            pts = rng.integers(-32, 16 * SIZE + 32, (3, 2)) / 16.0
        else:             # the 1/32 grid: exercises the snap itself (odd multiples of 1/32 are ties -> round to even)
            pts = rng.integers(-64, 32 * SIZE + 64, (3, 2)) / 32.0
        yield [(float(x), float(y)) for x, y in pts]


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_oracle_coverage_matches_the_independent_restatement(seed):
    checked = on_edge_cases = 0
    for pts in _random_triangles(80, seed):
        want = reference_coverage(pts)
        got = covered(Scene("x", SIZE, SIZE, tri(pts, size=(SIZE, SIZE))))
        assert np.array_equal(got, want), f"triangle {pts}: {np.argwhere(got != want)[:5].tolist()}"
        checked += int(want.sum())
        on_edge_cases += 1
    assert checked > 2000 and on_edge_cases == 80


def test_shared_edge_is_covered_exactly_once_by_the_restatement_too():
    # the property the top-left rule exists for, on the restatement itself: two triangles sharing an edge partition the quad
    a = reference_coverage([(2, 3), (28, 5), (4, 27)])
    b = reference_coverage([(28, 5), (30, 29), (4, 27)])
    assert not (a & b).any() and (a | b).sum() > 200


def test_perspective_correct_interpolation_against_float64():
    """Vertex colours are interpolated perspective-correctly (colour/w and 1/w screen-linear, then divided), depth z/w
    screen-linear.  A float64 restatement from the snapped vertices must agree with the oracle's fp32 pipeline to within
    1 LSB of UNORM8 / 1e-6 in depth on every covered pixel; with all w equal the result is plain barycentric."""
    import numpy as np
    from kazan_b200 import abi
    from tests import util
    size = 32
    clip = np.array([  # x, y, z, w, r, g, b, a in clip space: w differs per vertex
        [-0.9 * 1.0, -0.8 * 1.0, 0.2 * 1.0, 1.0, 1.0, 0.0, 0.0, 1.0],
        [0.8 * 3.0, -0.5 * 3.0, 0.6 * 3.0, 3.0, 0.0, 1.0, 0.0, 0.5],
        [-0.2 * 0.5, 0.9 * 0.5, 0.9 * 0.5, 0.5, 0.0, 0.0, 1.0, 0.25]], dtype=np.float32)
    sc = Scene("p", size, size, clip, pos_comps=4, color_offset=16, use_depth=True,
               state=util.pipeline_state(depth=(abi.COMPARE_LESS, 1)))
    color, depth, stats = sc.run_oracle()
    px = color.view(np.uint8)[:size * size * 4].reshape(size, size, 4)
    dz = depth.view(np.float32)[:size * size].reshape(size, size)
    w = clip[:, 3].astype(np.float64)
    ndc = clip[:, :3].astype(np.float64) / w[:, None]
    win = np.stack([(ndc[:, 0] * 0.5 + 0.5) * size, (ndc[:, 1] * 0.5 + 0.5) * size], axis=1)
    snapped = np.rint(win * 16.0) / 16.0  # the 1/16-pixel grid of the rasteriser
    (x0, y0), (x1, y1), (x2, y2) = snapped
    area = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0)
    checked = 0
    for y in range(size):
        for x in range(size):
            if dz[y, x] == 1.0:
                continue  # not covered (cleared depth)
            cx, cy = x + 0.5, y + 0.5
            l0 = ((x1 - cx) * (y2 - cy) - (x2 - cx) * (y1 - cy)) / area
            l1 = ((x2 - cx) * (y0 - cy) - (x0 - cx) * (y2 - cy)) / area
            l2 = 1.0 - l0 - l1
            lam = np.array([l0, l1, l2])
            assert lam.min() > -1e-9
            want_z = float(lam @ ndc[:, 2])
            assert abs(float(dz[y, x]) - want_z) < 1e-6, (x, y)
            inv_w = lam / w
            rgba = (inv_w @ clip[:, 4:8].astype(np.float64)) / inv_w.sum()
            want = np.clip(rgba, 0, 1) * 255.0
            assert np.all(np.abs(px[y, x].astype(np.float64) - want) <= 1.0 + 1e-6), (x, y, px[y, x], want)
            checked += 1
    assert checked == stats.fragments_covered > 150
    # the centroid of the screen triangle is NOT the average colour when w differs: perspective correction is really there
    gx, gy = int(snapped[:, 0].mean()), int(snapped[:, 1].mean())
    assert abs(int(px[gy, gx, 2]) - 85) > 40  # blue (the w = 0.5 vertex) dominates: 1/w weights 1 : 1/3 : 2


def test_blend_equations_against_float64():
    """Random blend states (10 factors x 5 ops for colour and alpha, write masks): a float64 restatement of the Vulkan
    "Blend Factors" / "Blend Operations" tables agrees with the oracle to within 1 LSB of UNORM8."""
    import numpy as np
    from tests import util
    rng = np.random.default_rng(77)

    def factor(f, s, d, ch):
        return [0.0, 1.0, s[ch], 1 - s[ch], d[ch], 1 - d[ch], s[3], 1 - s[3], d[3], 1 - d[3]][f]

    def op(o, cs, sf, cd, df):
        return [cs * sf + cd * df, cs * sf - cd * df, cd * df - cs * sf, min(cs, cd), max(cs, cd)][o]

    worst = 0.0
    for k in range(300):
        blend = tuple(int(v) for v in (rng.integers(0, 10), rng.integers(0, 10), rng.integers(0, 5),
                                       rng.integers(0, 10), rng.integers(0, 10), rng.integers(0, 5)))
        mask = int(rng.choice([15, 15, 7, 9, 6]))
        src = rng.uniform(-0.2, 1.2, 4)
        dst8 = rng.integers(0, 256, 4)
        t = tri([(0, 0), (8, 0), (0, 8)], color=[float(v) for v in src])
        sc = Scene("b", 8, 8, t, state=util.pipeline_state(blend=blend, write_mask=mask), clear_color=tuple(dst8 / 255.0))
        color, _, _ = sc.run_oracle()
        got = color[:4].astype(np.float64)
        s = np.clip(np.float32(src).astype(np.float64), 0, 1)
        d = dst8 / 255.0
        want = np.empty(4)
        for ch in range(3):
            want[ch] = op(blend[2], s[ch], factor(blend[0], s, d, ch), d[ch], factor(blend[1], s, d, ch))
        want[3] = op(blend[5], s[3], factor(blend[3], s, d, 3), d[3], factor(blend[4], s, d, 3))
        want = np.clip(want, 0, 1) * 255.0
        for ch in range(4):
            if mask >> ch & 1:
                err = abs(got[ch] - want[ch])
                worst = max(worst, err)
                assert err <= 1.0 + 1e-6, (k, blend, ch, got[ch], want[ch])
            else:
                assert got[ch] == dst8[ch], (k, "masked channel changed")
    assert worst <= 1.0 + 1e-6
