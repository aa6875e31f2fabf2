"""CPU: rule R9 of the oracle (points and lines as two triangles in snapped window coordinates) against a second derivation
straight from Vulkan 1.1 "Basic Point Rasterization" / "Basic Line Segment Rasterization" (non-strict lines, width 1, point
size 1 — the reference's limits, api_impl.rs:1604-1608), in exact integers:

* a point at snapped (X, Y) (1/16 pixel) produces exactly the pixel whose centre lies in [X-8, X+8) x [Y-8, Y+8);
* an x-major segment produces, for every pixel column whose centre x satisfies min(Xa, Xb) <= 16 px + 8 < max(Xa, Xb), exactly
  the one pixel whose centre lies within half a pixel of the line at that column (a centre exactly half a pixel away belongs
  to the side whose boundary is a top or a left edge) — and the same with x and y exchanged for y-major segments;
* every covered pixel of an additive draw is hit once (no double hits where the two triangles meet, none between the
  segments of a strip).
The reference has no known answers for rasterisation: like tests/test_oracle_crosscheck.py this pins the oracle
by deriving the result twice."""
from fractions import Fraction

import numpy as np
import pytest

from kazan_b200 import abi
from tests.util import Scene, pipeline_state

W, H = 64, 32  # powers of two: NDC -> window is exact in fp32 for positions on the 1/32-pixel grid


def ndc(px, py):
    return [px / W * 2 - 1, py / H * 2 - 1]


def run(verts, topology, blend=None):
    st = pipeline_state(topology=topology, blend=blend)
    color, _, stats = Scene("t", W, H, np.array(verts, dtype=np.float32), state=st).run_oracle()
    return color.view(np.uint32).reshape(H, W), stats


def snap(v):
    return int(np.rint(np.float32(v) * np.float32(16.0)))


def test_a_point_is_the_pixel_whose_centre_its_unit_square_holds():
    rng = np.random.default_rng(7)
    for _ in range(300):
        px, py = rng.integers(-32, W * 32 + 32) / 32.0, rng.integers(-32, H * 32 + 32) / 32.0
        img, stats = run([[*ndc(px, py), 0.5, 1, 1, 1, 1]], abi.TOPOLOGY_POINT_LIST)
        X, Y = snap(px), snap(py)
        want = {(x, y) for x in range(W) for y in range(H) if X - 8 <= 16 * x + 8 < X + 8 and Y - 8 <= 16 * y + 8 < Y + 8}
        got = {(int(x), int(y)) for y, x in zip(*np.nonzero(img))}
        assert got == want and len(want) <= 1, (px, py)
        assert stats.fragments_covered == len(want)


def line_pixels(Xa, Ya, Xb, Yb):
    """Second derivation, exact: the pixels of the non-strict parallelogram of a segment in 1/16-pixel coordinates.  A sample
    exactly on the parallelogram's outline belongs to it when that side is a top edge (horizontal, interior below) or a left
    edge (not horizontal, interior to its right at the same height) — Vulkan 1.1 "Basic Polygon Rasterization"."""
    out = set()
    if abs(Xb - Xa) >= abs(Yb - Ya):  # x-major: vertical ends, the two long sides half a pixel above and below the line
        if Xa == Xb:
            return out
        lo, hi = min(Xa, Xb), max(Xa, Xb)
        descending = (Yb - Ya) * (Xb - Xa) > 0  # y grows with x: the LOWER long side is the left edge, else the upper one (or it is a top edge)
        for px in range(-2, 200):
            cx = 16 * px + 8
            if not (lo <= cx < hi):  # the left end is a left edge, the right end a right edge
                continue
            yl = Fraction(Ya) + Fraction(Yb - Ya) * Fraction(cx - Xa, Xb - Xa)
            for py in range(-2, 100):
                cy = 16 * py + 8
                if (yl - 8 < cy <= yl + 8) if descending else (yl - 8 <= cy < yl + 8):
                    out.add((px, py))
    else:  # y-major: horizontal ends (top owned, bottom not), the long sides half a pixel left and right (the left one is a left edge)
        lo, hi = min(Ya, Yb), max(Ya, Yb)
        for py in range(-2, 100):
            cy = 16 * py + 8
            if not (lo <= cy < hi):
                continue
            xl = Fraction(Xa) + Fraction(Xb - Xa) * Fraction(cy - Ya, Yb - Ya)
            for px in range(-2, 200):
                cx = 16 * px + 8
                if xl - 8 <= cx < xl + 8:
                    out.add((px, py))
    return out


@pytest.mark.parametrize("seed", [1, 2])
def test_a_segment_is_the_non_strict_parallelogram(seed):
    rng = np.random.default_rng(seed)
    for k in range(250):
        ax, ay = rng.integers(0, W * 2) / 2.0, rng.integers(0, H * 2) / 2.0
        if k % 4 == 0:
            bx, by = ax + rng.integers(-20, 20), ay + rng.integers(-3, 3) / 2.0
        elif k % 4 == 1:
            bx, by = ax + rng.integers(-3, 3) / 2.0, ay + rng.integers(-20, 20)
        else:
            bx, by = rng.integers(0, W * 32) / 32.0, rng.integers(0, H * 32) / 32.0
        img, stats = run([[*ndc(ax, ay), 0.5, 1, 1, 1, 1], [*ndc(bx, by), 0.5, 1, 1, 1, 1]], abi.TOPOLOGY_LINE_LIST)
        want = {(x, y) for (x, y) in line_pixels(snap(ax), snap(ay), snap(bx), snap(by)) if 0 <= x < W and 0 <= y < H}
        got = {(int(x), int(y)) for y, x in zip(*np.nonzero(img))}
        assert got == want, (ax, ay, bx, by, sorted(got ^ want))
        assert stats.fragments_covered == len(want)


def test_strips_and_quads_hit_every_pixel_once():
    """Additive blending with colour 1/255 per hit: any pixel covered twice (the diagonal the two triangles share, the joint of
    two segments of a strip) would read 2."""
    rng = np.random.default_rng(3)
    one = 1.0 / 255.0
    additive = (1, 1, 0, 1, 1, 0)  # ONE, ONE, ADD
    for _ in range(40):
        n = int(rng.integers(2, 12))
        # a monotone polyline (x strictly increasing, slopes below 1): x-major segments that meet end to end
        xs = np.cumsum(rng.integers(1, 12, n)) / 2.0 + 1.0
        ys = np.clip(H / 2 + np.cumsum(rng.integers(-1, 2, n) * rng.integers(0, 2, n)) / 2.0, 1, H - 2)
        verts = [[*ndc(x, y), 0.5, one, one, one, one] for x, y in zip(xs, ys)]
        img, _ = run(verts, abi.TOPOLOGY_LINE_STRIP, blend=additive)
        assert int((img & 0xFF).max()) <= 1, "a pixel of a line strip was covered twice"
        cols = np.nonzero((img & 0xFF).sum(axis=0))[0]
        first, last = int(np.ceil((xs[0] * 16 - 8) / 16)), int(np.ceil((xs[-1] * 16 - 8) / 16)) - 1
        assert cols.tolist() == list(range(max(first, 0), min(last, W - 1) + 1)), "a column of the polyline has no pixel"


def test_line_attributes_run_along_the_major_axis_perspective_correct():
    """R9 / R6 against float64: at a covered pixel the colour is the perspective-correct blend of the two ends' colours at the
    pixel centre's position along the major axis, the depth the screen-linear one (within 1 LSB of UNORM8 / 1e-6)."""
    rng = np.random.default_rng(5)
    for k in range(60):
        ax, ay = rng.integers(2 * 2, (W - 2) * 2) / 2.0, rng.integers(2 * 2, (H - 2) * 2) / 2.0
        bx, by = rng.integers(2 * 2, (W - 2) * 2) / 2.0, rng.integers(2 * 2, (H - 2) * 2) / 2.0
        wa, wb = float(rng.uniform(0.5, 4.0)), float(rng.uniform(0.5, 4.0))
        za, zb = float(rng.uniform(0.0, 1.0)), float(rng.uniform(0.0, 1.0))
        ca, cb = rng.uniform(0, 1, 4), rng.uniform(0, 1, 4)
        va = [c * wa for c in ndc(ax, ay)] + [za * wa, wa, *ca]
        vb = [c * wb for c in ndc(bx, by)] + [zb * wb, wb, *cb]
        st = pipeline_state(topology=abi.TOPOLOGY_LINE_LIST, depth=(abi.COMPARE_ALWAYS, 1))
        sc = Scene("t", W, H, np.array([va, vb], dtype=np.float32), state=st, pos_comps=4, color_offset=16, use_depth=True, clear_depth=1.0)
        color, depth, _ = sc.run_oracle()
        img = color.reshape(H, W, 4)
        dep = depth.view(np.float32).reshape(H, W)
        Xa, Ya, Xb, Yb = snap(ax), snap(ay), snap(bx), snap(by)
        x_major = abs(Xb - Xa) >= abs(Yb - Ya)
        hits = np.argwhere(dep != 1.0)
        if Xa == Xb and Ya == Yb:
            assert len(hits) == 0
            continue
        for py, px in hits:
            t = (16 * px + 8 - Xa) / (Xb - Xa) if x_major else (16 * py + 8 - Ya) / (Yb - Ya)
            assert -1e-9 <= t <= 1 + 1e-9
            fwa, fwb = np.float64(np.float32(wa)), np.float64(np.float32(wb))
            den = (1 - t) / fwa + t / fwb
            want_c = ((1 - t) * np.float32(ca).astype(np.float64) / fwa + t * np.float32(cb).astype(np.float64) / fwb) / den
            got_c = img[py, px].astype(np.float64) / 255.0
            assert np.all(np.abs(got_c - want_c) <= 1.0 / 255.0 + 1e-6), (k, px, py, got_c, want_c)
            za32, zb32 = np.float64(np.float32(za * wa) / np.float32(wa)), np.float64(np.float32(zb * wb) / np.float32(wb))
            assert abs(float(dep[py, px]) - ((1 - t) * za32 + t * zb32)) <= 2e-6, (k, px, py)
