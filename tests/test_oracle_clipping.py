"""CPU: rule R3 of oracle/raster.c (clipping) checked against a derivation that does not clip.

The reference holds no known answers for rasterisation: the oracle defines what the GPU kernels must
produce, so the oracle's clipper is checked here against the mathematics it approximates.  A triangle in homogeneous clip
coordinates covers the pixel centre (xn, yn) of NDC space exactly where the solution l of

    [x0 x1 x2]       [xn]
    [y0 y1 y2] l  =  [yn]        (the point of the triangle's plane that projects to the pixel, scaled to w = 1)
    [w0 w1 w2]       [ 1]

has no negative component (Olano & Greer's homogeneous rasterisation: no clipping anywhere), and an attribute there is
sum(l_i a_i) / sum(l_i).  Evaluated in float64 per pixel, for triangles through the eye plane (w <= 0 vertices) and triangles millions
of pixels large — the ones the oracle clips.  Pixels closer than 0.2 px to an edge are skipped (the oracle snaps vertices to
1/16 px; a clipped polygon's new vertices are snapped too)."""
import numpy as np
import pytest

from kazan_b200 import abi
from tests import util
from tests.util import Scene, pipeline_state

W, H = 96, 64


def homogeneous(verts, px, py):
    """l (3,) for the pixel centre, float64; None when the triangle is edge-on"""
    xn = (px + 0.5) / (W / 2) - 1.0
    yn = (py + 0.5) / (H / 2) - 1.0
    m = np.array([[v[0] for v in verts], [v[1] for v in verts], [v[3] for v in verts]], dtype=np.float64)
    if abs(np.linalg.det(m)) < 1e-12:
        return None
    return np.linalg.solve(m, np.array([xn, yn, 1.0]))


def check(verts, min_checked):
    """verts: three rows x y z w r g b a (z = w/2: every fragment has depth 0.5, nothing is depth-clipped)"""
    sc = Scene("t", W, H, np.array(verts, dtype=np.float32), pos_comps=4, color_offset=16, state=pipeline_state())
    color, _, stats = sc.run_oracle()
    px = color.view(np.uint8)[:W * H * 4].reshape(H, W, 4)
    checked = 0
    for y in range(H):
        for x in range(W):
            ls = [homogeneous(verts, x + dx, y + dy) for dx, dy in ((0, 0), (-0.2, -0.2), (0.2, -0.2), (-0.2, 0.2), (0.2, 0.2))]
            if any(l is None for l in ls):
                continue
            inside = [bool((l >= 0).all()) for l in ls]
            if len(set(inside)) != 1:
                continue  # near an edge
            checked += 1
            got = px[y, x]
            if not inside[0]:
                assert (got == 0).all(), (x, y, got)
                continue
            want = sum(ls[0][i] * np.array(verts[i][4:8], dtype=np.float64) for i in range(3)) / ls[0].sum()
            want8 = np.clip(want, 0, 1) * 255.0
            assert np.abs(got.astype(np.float64) - want8).max() <= 1.01, (x, y, got, want8)
    assert checked >= min_checked
    return stats


def vert(x, y, w, rgba):
    return [x * 1.0, y * 1.0, 0.5 * w, w * 1.0, *rgba]


def test_triangle_through_the_eye_plane_is_clipped_not_dropped():
    # two vertices in front of the eye, one behind it: the visible part is an unbounded wedge of the screen
    verts = [vert(-0.6, 0.5, 1.0, (1, 0, 0, 1)), vert(0.7, 0.6, 2.0, (0, 1, 0, 1)), vert(0.1, 0.4, -1.5, (0, 0, 1, 1))]
    st = check(verts, 4000)
    assert st.triangles_clipped == 1 and st.clip_triangles >= 2 and st.triangles_culled == 0 and st.fragments_covered > 500


def test_triangle_with_two_vertices_behind_the_eye():
    verts = [vert(0.05, -0.1, 0.8, (1, 1, 0, 1)), vert(0.9, 0.3, -0.7, (0, 1, 1, 1)), vert(-0.8, 0.5, -2.0, (1, 0, 1, 1))]
    st = check(verts, 4000)
    assert st.triangles_clipped == 1 and st.fragments_covered > 200


def test_triangle_wholly_behind_the_eye_draws_nothing():
    verts = [vert(0.0, -0.5, -1.0, (1, 1, 1, 1)), vert(0.5, 0.5, -1.0, (1, 1, 1, 1)), vert(-0.5, 0.5, -2.0, (1, 1, 1, 1))]
    st = check(verts, 5000)
    assert st.triangles_clipped == 1 and st.clip_triangles == 0 and st.fragments_covered == 0 and st.triangles_culled == 1


@pytest.mark.parametrize("seed", range(6))
def test_triangles_far_larger_than_the_guard_band(seed):
    rng = np.random.default_rng(100 + seed)
    r = 10.0 ** rng.uniform(4.0, 6.5)  # NDC units: up to 10^8 pixels
    ang = rng.uniform(0, 2 * np.pi) + np.array([0.0, 2.0, 4.1])
    c = rng.uniform(-0.5, 0.5, 2)
    ws = rng.uniform(0.5, 3.0, 3)
    verts = []
    for i in range(3):
        far = 1.0 if i else rng.choice([1.0, 1e-5])  # sometimes one vertex on the screen
        verts.append(vert((c[0] + r * far * np.cos(ang[i])) * ws[i], (c[1] + r * far * np.sin(ang[i])) * ws[i], ws[i], rng.uniform(0, 1, 4)))
    st = check(verts, 4000)
    assert st.triangles_clipped == 1 and st.triangles_discarded == 0


def test_nan_and_infinity_are_discarded_and_counted():
    good = [vert(-0.5, -0.5, 1.0, (1, 1, 1, 1)), vert(0.5, -0.5, 1.0, (1, 1, 1, 1)), vert(0.0, 0.5, 1.0, (1, 1, 1, 1))]
    bad1 = [list(v) for v in good]
    bad1[1][0] = float("nan")
    bad2 = [list(v) for v in good]
    bad2[2][1] = float("inf")
    sc = Scene("t", W, H, np.array(bad1 + good + bad2, dtype=np.float32), pos_comps=4, color_offset=16, state=pipeline_state())
    _, _, st = sc.run_oracle()
    assert (st.triangles_in, st.triangles_discarded, st.triangles_clipped, st.triangles_culled) == (3, 2, 0, 2) and st.fragments_covered > 0


def test_floor_through_the_eye_plane_covers_every_pixel_once():
    """a floor of 12 x 16 quads from far in front of the eye to behind it, one colour, additive blending: below the horizon
    every pixel is hit exactly once — the triangles the clipper makes share their edges like any others"""
    near, far = 0.1, 100.0
    a, b = far / (far - near), -far * near / (far - near)
    verts = []
    nx, nz = 12, 16
    for j in range(nz + 1):
        for i in range(nx + 1):
            X = -30.0 + 60.0 * i / nx
            dist = 50.0 - 58.0 * j / nz  # 50 in front ... 8 behind
            verts.append([1.2 * X, 1.6, a * dist + b, dist, 0.25, 0.25, 0.25, 0.25])
    idx = []
    for j in range(nz):
        for i in range(nx):
            v00 = j * (nx + 1) + i
            idx += [v00, v00 + 1, v00 + nx + 2, v00, v00 + nx + 2, v00 + nx + 1]
    sc = Scene("floor", W, H, np.array(verts, dtype=np.float32), np.array(idx, dtype=np.uint32), pos_comps=4, color_offset=16,
               state=pipeline_state(blend=util.ADDITIVE_BLEND))
    color, _, st = sc.run_oracle()
    px = color.view(np.uint32)[:W * H].reshape(H, W)
    assert st.triangles_clipped > 0
    one_layer = px[H - 1, W // 2]
    assert one_layer == 0x40404040
    assert set(np.unique(px).tolist()) <= {0, int(one_layer)}, "a pixel was hit twice (or by a sliver the clipper made)"
    # y_ndc = 1.6 / w: the horizon is the middle row; the near plane (w = 0.1) is far below the frame, so everything below the
    # horizon belongs to the floor
    assert (px[H // 2 + 2:, :] == one_layer).all() and (px[:H // 2, :] == 0).all()
