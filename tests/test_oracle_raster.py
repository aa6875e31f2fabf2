"""CPU: the rasteriser oracle against hand-computed cases."""
import ctypes as C

import numpy as np
import pytest

from kazan_b200 import abi
from tests import cases, util
from tests.util import Scene, pipeline_state

W = [1, 1, 1, 1]


def covered(scene) -> np.ndarray:
    color, _, _ = scene.run_oracle()
    return color.view(np.uint32)[:scene.width * scene.height].reshape(scene.height, scene.width) != 0


def tri(points, color=W, z=0.5, size=(8, 8)):
    """Triangle given in PIXEL coordinates on a size[0] x size[1] target."""
    w, h = size
    return [[2.0 * x / w - 1.0, 2.0 * y / h - 1.0, z] + list(color) for x, y in points]


def test_top_left_rule_axis_aligned():
    # right triangle (0,0) (4,0) (0,4): hypotenuse x + y = 4 passes exactly through pixel centres with x+y == 3
    m = covered(Scene("t", 8, 8, tri([(0, 0), (4, 0), (0, 4)])))
    expect = np.zeros((8, 8), bool)
    for y in range(8):
        for x in range(8):
            expect[y, x] = (x + 0.5) + (y + 0.5) < 4  # centres ON the hypotenuse are out: not a top or left edge
    assert np.array_equal(m, expect)
    # a square from two triangles with edges through pixel centres: left/top edges own the centre, right/bottom do not
    sq = tri([(1.5, 1.5), (5.5, 1.5), (5.5, 5.5)]) + tri([(1.5, 1.5), (5.5, 5.5), (1.5, 5.5)])
    m = covered(Scene("s", 8, 8, sq))
    expect = np.zeros((8, 8), bool)
    expect[1:5, 1:5] = True
    assert np.array_equal(m, expect)


def test_winding_does_not_change_coverage():
    a = covered(Scene("a", 16, 16, tri([(1.3, 2.1), (14.2, 4.9), (6.6, 13.7)], size=(16, 16))))
    b = covered(Scene("b", 16, 16, tri([(1.3, 2.1), (6.6, 13.7), (14.2, 4.9)], size=(16, 16))))
    assert np.array_equal(a, b) and a.sum() > 20


def test_shared_edges_cover_every_pixel_exactly_once():
    for name in ("quad_shared_diagonal", "fan_subpixel_vertices"):
        sc = next(s for s in cases.draw_scenes() if s.name == name)
        color, _, stats = sc.run_oracle()
        px = color.view(np.uint32)[:sc.width * sc.height]
        assert set(np.unique(px)) <= {0, 0x40404040}, name  # 0.25 added once; twice would give 0x80
        assert stats.fragments_covered == int((px != 0).sum())
    sc = next(s for s in cases.draw_scenes() if s.name == "quad_shared_diagonal")
    assert sc.run_oracle()[2].fragments_covered == sc.width * sc.height


def test_degenerate_and_culled_triangles_produce_nothing():
    zero_area = tri([(1, 1), (4, 4), (7, 7)])
    _, _, st = Scene("z", 8, 8, zero_area).run_oracle()
    assert (st.triangles_in, st.triangles_culled, st.fragments_covered) == (1, 1, 0)
    t = tri([(0, 0), (8, 0), (0, 8)])  # 2A > 0: back facing under COUNTER_CLOCKWISE (Vulkan area a = -A)
    assert covered(Scene("b", 8, 8, t, state=pipeline_state(cull=abi.CULL_BACK))).sum() == 0
    assert covered(Scene("f", 8, 8, t, state=pipeline_state(cull=abi.CULL_FRONT))).sum() > 0
    assert covered(Scene("f", 8, 8, t, state=pipeline_state(cull=abi.CULL_FRONT, front=abi.FRONT_FACE_CW))).sum() == 0


def test_snapping_is_round_to_nearest_even_at_one_sixteenth():
    # vertex x = 3 + 1/32 px is a tie between 48/16 and 49/16 -> even (48): the edge x = 3.0 leaves centre 2.5 out
    # and, being a left edge, a vertex snapped to 2.5 owns centre 2.5
    for xv, first in ((3.0 + 1 / 32, 3), (2.5 + 1 / 32 + 1 / 16, 3), (2.5 + 1 / 64, 2)):
        m = covered(Scene("s", 8, 8, tri([(xv, 0), (8, 0), (xv, 8)]) + tri([(xv, 8), (8, 0), (8, 8)])))
        assert m[:, first].all() and not m[:, :first].any(), xv


def test_depth_test_and_write():
    near = tri([(0, 0), (8, 0), (0, 8)], color=[1, 0, 0, 1], z=0.25)
    far = tri([(0, 0), (8, 0), (0, 8)], color=[0, 1, 0, 1], z=0.75)
    st = pipeline_state(depth=(abi.COMPARE_LESS, 1))
    for order in (near + far, far + near):
        color, depth, _ = Scene("d", 8, 8, order, state=st, use_depth=True).run_oracle()
        px = color.view(np.uint32).reshape(8, 8)
        assert px[0, 0] == 0xFF0000FF  # red wins in both submission orders
        assert depth.view(np.float32).reshape(8, 8)[0, 0] == np.float32(0.25)
        assert depth.view(np.float32).reshape(8, 8)[7, 7] == np.float32(1.0)


def test_depth_interpolation_is_screen_linear():
    # z varies 0 -> 1 along x over 8 pixels: centre of pixel x has z = (x + 0.5) / 8 up to fp32 rounding of R6
    t = [[-1, -1, 0.0] + W, [1, -1, 1.0] + W, [-1, 1, 0.0] + W]
    _, depth, _ = Scene("z", 8, 8, t, state=pipeline_state(depth=(abi.COMPARE_ALWAYS, 1)), use_depth=True).run_oracle()
    row = depth.view(np.float32).reshape(8, 8)[0]
    assert np.allclose(row[:7], (np.arange(7) + 0.5) / 8, atol=2e-7)


def test_alpha_blend_known_answer():
    t = tri([(0, 0), (8, 0), (0, 8)], color=[1.0, 0.5, 0.0, 0.5])
    color, _, _ = Scene("b", 8, 8, t, state=pipeline_state(blend=util.ALPHA_BLEND), clear_color=(0, 0, 1, 1)).run_oracle()
    r, g, b, a = color[:4]
    # src*0.5 + dst*0.5 with dst = (0,0,1,1): (0.5, 0.25, 0.5, 0.75) -> RNE of x*255
    assert (r, g, b, a) == (128, 64, 128, 191)


def test_write_mask_and_no_blend_clamps_source():
    t = tri([(0, 0), (8, 0), (0, 8)], color=[2.0, -1.0, float("nan"), 0.5])
    color, _, _ = Scene("m", 8, 8, t, state=pipeline_state(write_mask=0b1011), clear_color=(0.2, 0.2, 0.2, 0.2)).run_oracle()
    assert tuple(color[:4]) == (255, 0, 51, 128)  # R clamped to 1, G clamped to 0, B masked (keeps 0.2), A = 0.5


def test_logic_ops_known_answers():
    """VkLogicOp on the UNORM8 bits: source (1, 0.2, 0.6, 0.4) -> FF 33 99 66, destination cleared to
    (0.0588.., 0.9411.., 0.3333.., 1) -> 0F F0 55 FF; blending is ignored when a logic op is enabled."""
    s, d = (0xFF, 0x33, 0x99, 0x66), (0x0F, 0xF0, 0x55, 0xFF)
    table = {0: lambda a, b: 0, 1: lambda a, b: a & b, 2: lambda a, b: a & ~b, 3: lambda a, b: a, 4: lambda a, b: ~a & b,
             5: lambda a, b: b, 6: lambda a, b: a ^ b, 7: lambda a, b: a | b, 8: lambda a, b: ~(a | b), 9: lambda a, b: ~(a ^ b),
             10: lambda a, b: ~b, 11: lambda a, b: a | ~b, 12: lambda a, b: ~a, 13: lambda a, b: ~a | b, 14: lambda a, b: ~(a & b),
             15: lambda a, b: 0xFF}
    t = tri([(0, 0), (8, 0), (0, 8)], color=[1.0, 0.2, 0.6, 0.4])
    clear = (15 / 255, 240 / 255, 85 / 255, 1.0)
    for op, fn in table.items():
        st = pipeline_state(blend=util.ALPHA_BLEND, logic_op=op)
        color, _, _ = Scene("l", 8, 8, t, state=st, clear_color=clear).run_oracle()
        assert tuple(color[:4]) == tuple(fn(a, b) & 0xFF for a, b in zip(s, d)), op
    # the write mask still applies: XOR on R and A only
    color, _, _ = Scene("l", 8, 8, t, state=pipeline_state(logic_op=6, write_mask=0b1001), clear_color=clear).run_oracle()
    assert tuple(color[:4]) == (0xFF ^ 0x0F, 0xF0, 0x55, 0x66 ^ 0xFF)


def test_strip_and_fan_assemble_the_triangles_of_the_specification():
    """Strip triangle i = {v[i], v[i+(1+i%2)], v[i+(2-i%2)]}, fan triangle i = {v[i+1], v[i+2], v[0]} (Vulkan 1.1 "Primitive
    Topologies"): the same picture as the equivalent triangle list, winding included (back faces culled)."""
    rng = np.random.default_rng(5)
    v = np.array([[x, y, 0.5, *rng.uniform(0.1, 0.9, 4)] for x, y in
                  [(-0.9, -0.8), (-0.7, 0.8), (-0.2, -0.7), (0.0, 0.9), (0.5, -0.9), (0.6, 0.7), (0.9, -0.6)]], dtype=np.float32)
    st = lambda topo, front=abi.FRONT_FACE_CCW: pipeline_state(topology=topo, cull=abi.CULL_BACK, front=front, blend=util.ADDITIVE_BLEND)
    n = len(v) - 2
    as_list = np.array([v[j] for i in range(n) for j in (i, i + (1 + i % 2), i + (2 - i % 2))], dtype=np.float32)
    strip, _, s_st = Scene("s", 64, 64, v, state=st(4)).run_oracle()
    lst, _, l_st = Scene("l", 64, 64, as_list, state=st(3)).run_oracle()
    assert strip.tobytes() == lst.tobytes() and s_st.triangles_in == n == l_st.triangles_in
    assert s_st.triangles_culled == l_st.triangles_culled == 0 and s_st.fragments_covered > 0  # one winding for all of them
    assert Scene("s", 64, 64, v, state=st(4, abi.FRONT_FACE_CW)).run_oracle()[2].triangles_culled == n
    fan_list = np.array([v[j] for i in range(n) for j in (i + 1, i + 2, 0)], dtype=np.float32)
    fan, _, f_st = Scene("f", 64, 64, v, state=pipeline_state(topology=5, blend=util.ADDITIVE_BLEND)).run_oracle()
    lst, _, l_st = Scene("l", 64, 64, fan_list, state=pipeline_state(blend=util.ADDITIVE_BLEND)).run_oracle()
    assert fan.tobytes() == lst.tobytes() and f_st.fragments_covered == l_st.fragments_covered > 0
    # fewer than three vertices: nothing; points and lines are not rasterised
    assert Scene("e", 8, 8, v[:2], state=st(4)).run_oracle()[2].triangles_in == 0


def test_scissor_viewport_and_band_limit_fragments():
    full = util.FULLSCREEN_TRIANGLE
    m = covered(Scene("s", 32, 32, full, scissor=abi.Rect2D(abi.Offset2D(3, 5), abi.Extent2D(7, 11))))
    e = np.zeros((32, 32), bool); e[5:16, 3:10] = True
    assert np.array_equal(m, e)
    m = covered(Scene("v", 32, 32, full, viewport=abi.Viewport(8, 4, 16, 20, 0, 1)))
    e = np.zeros((32, 32), bool); e[4:24, 8:24] = True
    assert np.array_equal(m, e)
    m = covered(Scene("b", 32, 32, full, band=(10, 14)))
    e = np.zeros((32, 32), bool); e[10:14] = True
    assert np.array_equal(m, e)


def test_z_clip_discards_fragments_outside_the_depth_range():
    t = [[-1, -1, -0.5] + W, [1, -1, 1.5] + W, [-1, 1, -0.5] + W]  # z from -0.5 to 1.5 across x
    m = covered(Scene("z", 16, 16, t))
    xs = np.where(m[0])[0]
    assert xs.min() == 4 and xs.max() <= 11  # z(x) = -0.5 + 2*(x+0.5)/16 in [0,1]  <=>  3.5 <= x <= 11.5


def test_procedural_fragment_program_matches_a_float32_numpy_restatement(oracle):
    Wd, Hd = 64, 36
    out = np.zeros(4, dtype=np.float32)
    f = np.float32
    for (px, py) in [(0, 0), (17, 9), (40, 18), (63, 35), (33, 3)]:
        oracle.kzo_fs_procedural_orbit(px, py, C.c_float(Wd), C.c_float(Hd), out.ctypes.data_as(C.POINTER(C.c_float)))
        u = (f(px) + f(0.5)) * (f(1) / f(Wd)); v = (f(py) + f(0.5)) * (f(1) / f(Hd))
        fma = lambda a, b, c: f(np.float64(a) * np.float64(b) + np.float64(c))  # exact product, one rounding
        cx, cy = fma(u, f(3.5), f(-2.5)), fma(v, f(2.0), f(-1.0))
        zx = zy = f(0)
        with np.errstate(all="ignore"):
            for _ in range(32):
                t = fma(zx, zx, cx); nzx = fma(-zy, zy, t); nzy = fma(zx + zx, zy, cy); zx, zy = nzx, nzy
            exp = [f(0.25) * fma(zx, zx, zy * zy), fma(zx, f(0.5), f(0.5)), fma(zy, f(0.5), f(0.5)), f(1)]
        for a, b in zip(out, exp):
            assert (np.isnan(a) and np.isnan(b)) or a == b, (px, py, out, exp)


def test_invalid_draw_state_is_rejected():
    sc = Scene("x", 8, 8, tri([(0, 0), (8, 0), (0, 8)]))
    sc.state.fragment_program = 7
    lib = __import__("oracle.binding", fromlist=["x"]).load()
    color, depth = sc._initial()
    fb = sc._framebuffer(color.ctypes.data, depth.ctypes.data)
    vi = abi.VertexInput(sc.vertices.ctypes.data, sc.vertices.nbytes, 28, 0, 3, 12)
    args = lambda st, vp: lib.kzo_draw(C.byref(st), C.byref(fb), C.byref(vp), C.byref(sc.scissor), C.byref(vi), 3, 1, 0, 0, None)
    assert args(sc.state, sc.viewport) == abi.ERROR_FEATURE_NOT_PRESENT
    wide = abi.VertexInput(sc.vertices.ctypes.data, sc.vertices.nbytes, 1 << 28, 0, 3, 12)
    assert lib.kzo_draw(C.byref(pipeline_state()), C.byref(fb), C.byref(sc.viewport), C.byref(sc.scissor), C.byref(wide), 3, 1, 0, 0,
                        None) == abi.ERROR_VALIDATION_FAILED  # index x stride must stay far from 2^64
    adjacency = pipeline_state(topology=6)  # LINE_LIST_WITH_ADJACENCY .. PATCH_LIST: not rasterised (points, lines and strips are: R9)
    assert args(adjacency, sc.viewport) == abi.ERROR_FEATURE_NOT_PRESENT
    assert args(pipeline_state(topology=11), sc.viewport) == abi.ERROR_VALIDATION_FAILED
    bad = pipeline_state()
    bad.logic_op = 17
    assert args(bad, sc.viewport) == abi.ERROR_VALIDATION_FAILED
    assert args(pipeline_state(), abi.Viewport(0, 0, 0, 8, 0, 1)) == abi.ERROR_VALIDATION_FAILED
    assert args(pipeline_state(), abi.Viewport(0, 0, 8, 8, 0, 2)) == abi.ERROR_VALIDATION_FAILED
    for bad_vp in (abi.Viewport(float("nan"), 0, 8, 8, 0, 1), abi.Viewport(0, float("inf"), 8, 8, 0, 1),
                   abi.Viewport(0, 0, float("inf"), 8, 0, 1), abi.Viewport(3e9, 0, 8, 8, 0, 1), abi.Viewport(0, 0, 8, -3e9, 0, 1)):
        assert args(pipeline_state(), bad_vp) == abi.ERROR_VALIDATION_FAILED  # outside viewportBoundsRange
