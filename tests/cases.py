"""Case tables shared by the oracle tests (CPU), the GPU parity tests and the golden generator.

Transfer cases follow SURVEY.md §8d config #1 and its stress variants; draw scenes are small versions of
configs #3/#4 plus the edge cases of SURVEY.md Appendix E (top-left rule, degenerate input, clipping by
scissor/viewport, depth ops, blend ops, tiled attachments, bands).
"""
from __future__ import annotations

import numpy as np

from kazan_b200 import abi
from kazan_b200.host import buffer_image_copy
from tests import util
from tests.util import Scene, pipeline_state

L, T = abi.TILING_LINEAR, abi.TILING_TILED

# (name, image dict, regions, buffer_bytes)
BUFFER_IMAGE_CASES = [
    ("cfg1_1024", dict(width=1024, height=1024), [buffer_image_copy(1024, 1024)], 1024 * 1024 * 4),
    ("padded_pitch_1056", dict(width=1024, height=256), [buffer_image_copy(1024, 256, row_length=1056)], 1056 * 256 * 4),
    ("subrect_16_8", dict(width=1024, height=1024),
     [buffer_image_copy(992, 1000, offset=(16, 8, 0))], 992 * 1000 * 4),
    ("misaligned_offset_4", dict(width=256, height=64), [buffer_image_copy(256, 64, buffer_offset=4)], 256 * 64 * 4 + 4),
    # source 4 / 8 / 12 bytes off a 16-byte boundary, destination aligned: the realigning copy (transfer.cu
    # copy_realign_kernel); 257 x 60 texels = 3855 vectors: a partial last warp and a partial tail vector
    ("misaligned_offset_8", dict(width=256, height=64), [buffer_image_copy(256, 64, buffer_offset=8)], 256 * 64 * 4 + 8),
    ("misaligned_offset_12_ragged", dict(width=257 * 4, height=15), [buffer_image_copy(257 * 4, 15, buffer_offset=12)], 257 * 4 * 15 * 4 + 12),
    ("odd_width_13", dict(width=13, height=7), [buffer_image_copy(13, 7)], 13 * 7 * 4),
    ("odd_subrect", dict(width=67, height=33), [buffer_image_copy(31, 17, offset=(5, 3, 0), row_length=40, image_height=20)],
     40 * 20 * 4),
    ("two_regions", dict(width=128, height=128),
     [buffer_image_copy(64, 128), buffer_image_copy(64, 128, buffer_offset=64 * 128 * 4, offset=(64, 0, 0))],
     128 * 128 * 4),
    ("array_layers", dict(width=64, height=32, array_layers=5),
     [buffer_image_copy(64, 32, base_layer=1, layer_count=3, image_height=40)], 64 * 40 * 3 * 4),
    ("volume_3d", dict(width=32, height=16, depth=6), [buffer_image_copy(20, 10, depth=4, offset=(4, 2, 1))], 20 * 10 * 4 * 4),
    ("single_texel", dict(width=9, height=9), [buffer_image_copy(1, 1, offset=(8, 8, 0))], 4),
    ("empty_extent", dict(width=16, height=16), [buffer_image_copy(0, 16)], 64),
    ("tiled_full", dict(width=128, height=64, tiling=T), [buffer_image_copy(128, 64)], 128 * 64 * 4),
    ("tiled_odd", dict(width=50, height=37, tiling=T), [buffer_image_copy(50, 37)], 50 * 37 * 4),
    ("tiled_subrect", dict(width=100, height=70, tiling=T, array_layers=2),
     [buffer_image_copy(41, 23, offset=(7, 19, 0), base_layer=1, row_length=48)], 48 * 23 * 4),
    # a single row into a TILED image: once took the contiguous-run shortcut of box_kernel and landed at texel (0, 0)
    # (found by the random differential test in tests/test_gpu_transfer.py)
    ("tiled_single_row", dict(width=70, height=41, tiling=T, array_layers=2),
     [buffer_image_copy(64, 1, buffer_offset=8, row_length=64, offset=(6, 3, 0), base_layer=1)], 8 + 64 * 4),
    ("tiled_aligned_subrect", dict(width=128, height=64, tiling=T), [buffer_image_copy(64, 32, offset=(32, 16, 0))],
     64 * 32 * 4),
    ("depth_format", dict(width=96, height=40, format=abi.FORMAT_D32_SFLOAT),
     [buffer_image_copy(96, 40, aspect=abi.ASPECT_DEPTH)], 96 * 40 * 4),
]

# (name, image dict, clear ranges) — ranges as (base_layer, layer_count)
CLEAR_CASES = [
    ("cfg1_1024", dict(width=1024, height=1024), [(0, abi.REMAINING)]),
    ("odd_3x5", dict(width=3, height=5), [(0, 1)]),
    ("layers_1_2_of_4", dict(width=33, height=17, array_layers=4), [(1, 2)]),
    ("two_ranges", dict(width=16, height=16, array_layers=4), [(0, 1), (3, 1)]),
    ("tiled_50x37", dict(width=50, height=37, tiling=T), [(0, abi.REMAINING)]),
    ("volume", dict(width=20, height=10, depth=3), [(0, 1)]),
]

CLEAR_COLOR = (0.25, 0.5, 0.75, 1.0)  # -> bytes 40 80 BF FF


def image_desc_fields(spec: dict) -> dict:
    d = dict(width=1, height=1, depth=1, array_layers=1, format=abi.FORMAT_R8G8B8A8_UNORM, tiling=L)
    d.update(spec)
    return d


# (name, src image, dst image, regions, filter)
def blit_cases():
    B = util.blit_region
    return [
        ("nearest_identity", dict(width=64, height=48), dict(width=64, height=48), [B((0, 0, 0), (64, 48, 1), (0, 0, 0), (64, 48, 1))],
         abi.FILTER_NEAREST),
        ("nearest_up2", dict(width=32, height=24), dict(width=64, height=48), [B((0, 0, 0), (32, 24, 1), (0, 0, 0), (64, 48, 1))],
         abi.FILTER_NEAREST),
        ("nearest_down3_mirror_x", dict(width=96, height=60), dict(width=40, height=40),
         [B((96, 0, 0), (0, 60, 1), (4, 10, 0), (36, 30, 1))], abi.FILTER_NEAREST),
        ("linear_up", dict(width=17, height=13), dict(width=64, height=48), [B((0, 0, 0), (17, 13, 1), (0, 0, 0), (64, 48, 1))],
         abi.FILTER_LINEAR),
        ("linear_down_mirror_y", dict(width=80, height=50), dict(width=33, height=21),
         [B((3, 2, 0), (77, 47, 1), (1, 20, 0), (32, 0, 1))], abi.FILTER_LINEAR),
        ("linear_tiled_src", dict(width=50, height=37, tiling=T), dict(width=70, height=50),
         [B((0, 0, 0), (50, 37, 1), (0, 0, 0), (70, 50, 1))], abi.FILTER_LINEAR),
        ("nearest_tiled_dst_layers", dict(width=20, height=20, array_layers=2), dict(width=45, height=31, tiling=T, array_layers=3),
         [B((0, 0, 0), (20, 20, 1), (5, 1, 0), (45, 31, 1), src_layer=0, dst_layer=1, layers=2)], abi.FILTER_NEAREST),
        # same size, not mirrored, NEAREST: the library turns these into box copies (16-byte vectors / 4-byte texels, TILED sides,
        # layers, 3D slices); the mirrored twin of the first one stays a blit
        ("nearest_same_size_offsets_tiled_src", dict(width=80, height=60, tiling=T), dict(width=96, height=70),
         [B((5, 3, 0), (69, 51, 1), (12, 9, 0), (76, 57, 1))], abi.FILTER_NEAREST),
        ("nearest_same_size_mirrored", dict(width=80, height=60, tiling=T), dict(width=96, height=70),
         [B((69, 3, 0), (5, 51, 1), (12, 9, 0), (76, 57, 1))], abi.FILTER_NEAREST),
        ("nearest_same_size_aligned_layers", dict(width=64, height=40, array_layers=3), dict(width=128, height=48, tiling=T, array_layers=2),
         [B((16, 8, 0), (48, 40, 1), (32, 16, 0), (64, 48, 1), src_layer=1, dst_layer=0, layers=2)], abi.FILTER_NEAREST),
        ("nearest_same_size_3d_slices", dict(width=24, height=16, depth=4), dict(width=24, height=16, depth=8),
         [B((0, 0, 0), (24, 16, 4), (0, 0, 0), (24, 16, 8))], abi.FILTER_NEAREST),
    ]


def copy_image_cases():
    Cp = util.image_copy
    return [
        ("linear_to_linear", dict(width=64, height=32), dict(width=80, height=40), [Cp((40, 20, 1), (3, 5, 0), (17, 9, 0))]),
        ("linear_to_tiled", dict(width=50, height=37), dict(width=50, height=37, tiling=T), [Cp((50, 37, 1))]),
        ("tiled_to_tiled_layers", dict(width=48, height=32, tiling=T, array_layers=3), dict(width=64, height=64, tiling=T, array_layers=3),
         [Cp((32, 16, 1), (16, 16, 0), (4, 32, 0), src_layer=1, dst_layer=0, layers=2)]),
        ("volume", dict(width=16, height=16, depth=4), dict(width=16, height=16, depth=4), [Cp((8, 8, 3), (0, 0, 1), (8, 8, 0))]),
        ("depth_to_depth", dict(width=31, height=9, format=abi.FORMAT_D32_SFLOAT), dict(width=31, height=9, format=abi.FORMAT_D32_SFLOAT),
         [Cp((31, 9, 1), aspect=abi.ASPECT_DEPTH)]),
    ]


CONVERT_TILING_CASES = [
    ("50x37", dict(width=50, height=37)), ("128x64_layers2", dict(width=128, height=64, array_layers=2)),
    ("1x1", dict(width=1, height=1)), ("16x16", dict(width=16, height=16)), ("17x33_vol2", dict(width=17, height=33, depth=2)),
]


# ---- draw scenes ------------------------------------------------------------------------------------------------

def _quad(color, z=0.5, x0=-1.0, y0=-1.0, x1=1.0, y1=1.0):
    c = list(color)
    return [[x0, y0, z] + c, [x1, y0, z] + c, [x1, y1, z] + c, [x0, y0, z] + c, [x1, y1, z] + c, [x0, y1, z] + c]


def draw_scenes() -> list[Scene]:
    rng = np.random.default_rng(20261017)
    scenes: list[Scene] = []
    # config #3 in small: fullscreen triangle + procedural fragment program
    scenes.append(Scene("fill_orbit_256x144", 256, 144, util.FULLSCREEN_TRIANGLE,
                        state=pipeline_state(fs=abi.FS_PROCEDURAL_ORBIT)))
    scenes.append(Scene("fill_orbit_odd_131x77_tiled", 131, 77, util.FULLSCREEN_TRIANGLE,
                        state=pipeline_state(fs=abi.FS_PROCEDURAL_ORBIT), color_tiling=T))
    # shared edges: every pixel exactly once (additive blend makes double hits visible)
    q = _quad((0.25, 0.25, 0.25, 0.25))
    scenes.append(Scene("quad_shared_diagonal", 64, 48, q, state=pipeline_state(blend=util.ADDITIVE_BLEND)))
    ctr = [0.013, -0.027, 0.5, 0.25, 0.25, 0.25, 0.25]
    ang = np.sort(rng.uniform(0, 2 * np.pi, 11))
    pts = [[float(np.cos(a)) * 1.7 + 0.01, float(np.sin(a)) * 1.9, 0.5, 0.25, 0.25, 0.25, 0.25] for a in ang]
    fan = []
    for i in range(len(pts)):
        fan += [ctr, pts[i], pts[(i + 1) % len(pts)]]
    scenes.append(Scene("fan_subpixel_vertices", 97, 61, fan, state=pipeline_state(blend=util.ADDITIVE_BLEND)))
    # config #4 in small: jittered grid mesh, D32 LESS + alpha blend (binned path: > 256 triangles)
    v, i = util.grid_mesh(40, 20, 4)
    scenes.append(Scene("mesh_40x20_depth_blend", 320, 180, v, i, state=pipeline_state(depth=(abi.COMPARE_LESS, 1), blend=util.ALPHA_BLEND),
                        use_depth=True))
    v, i = util.grid_mesh(200, 100, 5)
    scenes.append(Scene("mesh_200x100_depth_blend", 480, 270, v, i, state=pipeline_state(depth=(abi.COMPARE_LESS, 1), blend=util.ALPHA_BLEND),
                        use_depth=True))
    v, i = util.grid_mesh(24, 12, 6)
    scenes.append(Scene("mesh_cull_back_cw", 200, 120, v, i, state=pipeline_state(cull=abi.CULL_BACK, front=abi.FRONT_FACE_CW)))
    scenes.append(Scene("mesh_cull_front_ccw_tiled", 200, 120, v, i, state=pipeline_state(cull=abi.CULL_FRONT), color_tiling=T))
    # viewport / scissor clipping, negative-height viewport, band
    v, i = util.grid_mesh(30, 30, 7)
    scenes.append(Scene("viewport_offset_scissor", 160, 120, v, i, viewport=abi.Viewport(21.5, 10.25, 100.0, 80.5, 0.1, 0.9),
                        scissor=abi.Rect2D(abi.Offset2D(30, 20), abi.Extent2D(77, 55)),
                        state=pipeline_state(depth=(abi.COMPARE_LESS_OR_EQUAL, 1)), use_depth=True))
    scenes.append(Scene("viewport_negative_height", 128, 96, v, i, viewport=abi.Viewport(0, 96, 128, -96, 0, 1),
                        state=pipeline_state(blend=util.ALPHA_BLEND)))
    scenes.append(Scene("band_rows_40_72", 128, 96, v, i, band=(40, 72), state=pipeline_state(blend=util.ALPHA_BLEND)))
    # index variants
    v16, i32 = util.grid_mesh(12, 9, 8)
    pad = np.zeros((5, 7), dtype=np.float32)
    scenes.append(Scene("uint16_first_index_vertex_offset", 90, 70, np.concatenate([pad, v16]), i32.astype(np.uint16),
                        index_type=abi.INDEX_TYPE_UINT16, first=6, vertex_offset=5))
    scenes.append(Scene("non_indexed_instances2", 64, 48, _quad((0.125, 0.25, 0.5, 0.125)), instances=2,
                        state=pipeline_state(blend=util.ADDITIVE_BLEND)))
    # perspective: vec4 positions with varying w
    pv = []
    for k in range(60):
        w = rng.uniform(0.5, 4.0, 3)
        xy = rng.uniform(-1.2, 1.2, (3, 2))
        z = rng.uniform(0, 1, 3)
        col = rng.uniform(0, 1, (3, 4))
        for j in range(3):
            pv.append([xy[j, 0] * w[j], xy[j, 1] * w[j], z[j] * w[j], w[j], *col[j]])
    scenes.append(Scene("perspective_w", 150, 110, np.array(pv, dtype=np.float32), pos_comps=4, color_offset=16,
                        state=pipeline_state(depth=(abi.COMPARE_LESS, 1)), use_depth=True))
    # degenerate / invalid input mixed with valid triangles
    bad = np.array(pv[:30], dtype=np.float32)
    bad[0, 3] = 0.0           # w == 0
    bad[4, 3] = -1.0          # w < 0
    bad[6, 0] = np.nan        # NaN position
    bad[9, 1] = np.inf
    bad[12:15, :4] = bad[12, :4]  # zero area
    bad[15, 0] = 3.0e7        # beyond the guard band
    bad[18:21, 2] = 2.0 * bad[18:21, 3]  # z > w: clipped away per fragment
    scenes.append(Scene("degenerate_mix", 150, 110, bad, pos_comps=4, color_offset=16,
                        state=pipeline_state(blend=util.ALPHA_BLEND)))
    # out-of-range indices fetch zeros
    vo, io = util.grid_mesh(6, 6, 9)
    io = io.copy()
    io[5] = 1000
    io[20] = 0xFFFFFFF0
    scenes.append(Scene("robust_index_fetch", 64, 64, vo, io))
    # depth ops and masks
    v, i = util.grid_mesh(20, 20, 10)
    scenes.append(Scene("depth_greater_nowrite_mask_rb", 100, 100, v, i, clear_depth=0.5, use_depth=True,
                        state=pipeline_state(depth=(abi.COMPARE_GREATER, 0), write_mask=5), clear_color=(0.2, 0.4, 0.6, 0.8)))
    scenes.append(Scene("depth_always_write", 100, 100, v, i, use_depth=True, state=pipeline_state(depth=(abi.COMPARE_ALWAYS, 1))))
    scenes.append(Scene("depth_never", 100, 100, v, i, use_depth=True, state=pipeline_state(depth=(abi.COMPARE_NEVER, 1))))
    scenes.append(Scene("depth_not_equal_blend_minmax", 100, 100, v, i, use_depth=True, clear_depth=0.5, clear_color=(0.5, 0.5, 0.5, 0.5),
                        state=pipeline_state(depth=(abi.COMPARE_NOT_EQUAL, 1),
                                             blend=(abi.BLEND_ONE, abi.BLEND_ONE, abi.BLEND_OP_MIN, abi.BLEND_ONE, abi.BLEND_ONE, abi.BLEND_OP_MAX))))
    scenes.append(Scene("blend_dst_factors_revsub", 100, 100, v, i, clear_color=(0.9, 0.6, 0.3, 0.7),
                        state=pipeline_state(blend=(abi.BLEND_DST_ALPHA, abi.BLEND_ONE_MINUS_SRC_COLOR, abi.BLEND_OP_REVERSE_SUBTRACT,
                                                    abi.BLEND_ONE_MINUS_DST_ALPHA, abi.BLEND_DST_COLOR, abi.BLEND_OP_SUBTRACT))))
    # one tile hammered by more triangles than the in-shared-memory sort holds (binned-list overflow path)
    many = []
    for k in range(4500):
        x0, y0 = rng.uniform(-1.0, 0.2, 2)
        many += [[x0, y0, 0.5, 0.01, 0.02, 0.03, 0.004], [x0 + 0.9, y0 + 0.1, 0.5, 0.01, 0.02, 0.03, 0.004],
                 [x0 + 0.2, y0 + 0.8, 0.5, 0.01, 0.02, 0.03, 0.004]]
    scenes.append(Scene("overflow_4500_tris_one_tile", 40, 24, np.array(many, dtype=np.float32),
                        state=pipeline_state(blend=util.ADDITIVE_BLEND)))
    # thin slivers and large triangles mixed, > 256 triangles so that wide triangles go through binning
    sl = []
    for k in range(300):
        a = rng.uniform(-1.3, 1.3, 2)
        d = rng.uniform(-1, 1, 2) * (2.0 if k % 7 == 0 else 0.3)
        n = np.array([-d[1], d[0]]) * (0.5 if k % 11 == 0 else 0.004)
        col = rng.uniform(0, 1, 4)
        zz = rng.uniform(0, 1)
        for pnt in (a, a + d, a + d + n):
            sl.append([pnt[0], pnt[1], zz, *col])
    scenes.append(Scene("slivers_and_large", 256, 160, np.array(sl, dtype=np.float32),
                        state=pipeline_state(depth=(abi.COMPARE_LESS, 1), blend=util.ALPHA_BLEND), use_depth=True))
    # logic ops (logicOpEnable): order-dependent XOR over the binned path, masked OR_REVERSE over the direct path (few
    # triangles), both on top of a non-trivial clear colour; blending state is set but must be ignored
    scenes.append(Scene("logic_xor_slivers", 256, 160, np.array(sl, dtype=np.float32),
                        state=pipeline_state(blend=util.ALPHA_BLEND, logic_op=6), clear_color=(0.3, 0.6, 0.9, 0.5)))
    scenes.append(Scene("logic_or_reverse_masked_quad", 64, 48, _quad((0.7, 0.1, 0.55, 0.35)),
                        state=pipeline_state(logic_op=11, write_mask=0b0111), clear_color=(0.2, 0.8, 0.4, 0.6)))
    scenes.append(Scene("logic_and_inverted_mesh_tiled", 200, 120, *util.grid_mesh(30, 16, 7),
                        state=pipeline_state(logic_op=4, depth=(abi.COMPARE_LESS_OR_EQUAL, 1)), use_depth=True,
                        clear_color=(0.9, 0.5, 0.1, 1.0), color_tiling=T))
    # triangle strips and fans (Vulkan "Primitive Topologies"): a ribbon whose odd triangles would be culled if the strip
    # were assembled as {i, i+1, i+2}; an indexed (uint16) strip over the binned path; a fan around an off-centre hub
    rib = []
    for k in range(24):
        x = -0.95 + 1.9 * k / 23
        rib.append([x, -0.5 + 0.1 * np.sin(k), 0.5, *rng.uniform(0.1, 0.9, 4)])
        rib.append([x + 0.03, 0.55 + 0.1 * np.cos(k), 0.5, *rng.uniform(0.1, 0.9, 4)])
    scenes.append(Scene("strip_ribbon_cull_back", 160, 96, np.array(rib, dtype=np.float32),
                        state=pipeline_state(topology=4, cull=abi.CULL_BACK, front=abi.FRONT_FACE_CCW, blend=util.ADDITIVE_BLEND)))
    gv, _ = util.grid_mesh(24, 12, 9)
    strip_idx = []
    for row in range(12):  # one long strip, rows joined by degenerate triangles
        for col in range(25):
            strip_idx += [row * 25 + col, (row + 1) * 25 + col]
        strip_idx += [(row + 1) * 25 + 24, (row + 1) * 25]
    scenes.append(Scene("strip_indexed_u16_mesh", 240, 140, gv, np.array(strip_idx, dtype=np.uint16), index_type=abi.INDEX_TYPE_UINT16,
                        state=pipeline_state(topology=4, depth=(abi.COMPARE_LESS, 1), blend=util.ALPHA_BLEND), use_depth=True))
    hub = [[0.21, -0.13, 0.4, 1.0, 1.0, 1.0, 0.5]]
    rim = [[0.21 + 0.9 * float(np.cos(a)), -0.13 + 0.8 * float(np.sin(a)), 0.6, *rng.uniform(0, 1, 4)]
           for a in np.linspace(0, 2 * np.pi, 41)]
    scenes.append(Scene("fan_disc", 128, 128, np.array(hub + rim, dtype=np.float32),
                        state=pipeline_state(topology=5, blend=util.ADDITIVE_BLEND)))
    # wide spans (>= 64 px of a row: quad path) interleaved with small triangles, order-dependent state
    mx = []
    for k in range(400):
        c = rng.uniform(-1.1, 1.1, 2)
        size = 1.6 if k % 5 == 0 else (0.5 if k % 3 == 0 else 0.05)
        col = rng.uniform(0, 1, 4)
        for _ in range(3):
            o = rng.uniform(-size, size, 2)
            mx.append([c[0] + o[0], c[1] + o[1], rng.uniform(0, 1), *col])
    scenes.append(Scene("mixed_sizes_blend_depth", 300, 170, np.array(mx, dtype=np.float32),
                        state=pipeline_state(depth=(abi.COMPARE_LESS_OR_EQUAL, 1), blend=util.ALPHA_BLEND), use_depth=True))
    scenes.append(Scene("mixed_sizes_additive_tiled", 300, 170, np.array(mx[:600], dtype=np.float32) * np.array([1, 1, 1, 0.1, 0.1, 0.1, 0.1], dtype=np.float32),
                        state=pipeline_state(blend=util.ADDITIVE_BLEND), color_tiling=T))
    scenes += clip_scenes(rng)
    return scenes


def _ground_plane(rng, nx, nz, x_half, z_far, z_behind, height=-1.0):
    """a jittered grid in the plane y = height of eye space, from z_far in front of the eye to z_behind BEHIND it, in clip
    coordinates of a perspective projection (w = distance in front of the eye, negative behind it): the rows of quads around
    z = 0 cross the eye plane.  Returns vertices (x y z w r g b a) and uint32 indices."""
    near, far = 0.1, 100.0
    a, b = far / (far - near), -far * near / (far - near)
    verts = []
    for j in range(nz + 1):
        for i in range(nx + 1):
            X = -x_half + 2 * x_half * i / nx + rng.uniform(-0.2, 0.2) * x_half / nx
            dist = z_far + (z_behind - z_far) * j / nz + rng.uniform(-0.2, 0.2) * (z_far - z_behind) / nz  # in front of the eye > 0
            verts.append([1.2 * X, -1.6 * height, a * dist + b, dist, *rng.uniform(0.05, 1.0, 4)])
    idx = []
    for j in range(nz):
        for i in range(nx):
            v00 = j * (nx + 1) + i
            idx += [v00, v00 + 1, v00 + nx + 2, v00, v00 + nx + 2, v00 + nx + 1]
    return np.array(verts, dtype=np.float32), np.array(idx, dtype=np.uint32)


def clip_scenes(rng) -> list[Scene]:
    """R3: triangles that cross the eye plane (w = 0) or leave the guard band are clipped, not dropped."""
    scenes: list[Scene] = []
    # a floor under the eye, two triangles (fused direct path), both crossing the eye plane
    fv, fi = _ground_plane(rng, 1, 1, 4.0, 30.0, -14.0)
    scenes.append(Scene("clip_floor_quad_eye_plane", 160, 100, fv, fi, pos_comps=4, color_offset=16,
                        state=pipeline_state(depth=(abi.COMPARE_LESS, 1)), use_depth=True))
    # the same floor finer: 8 x 12 quads = 192 triangles (direct path with setup_kernel), additive blending shows every
    # pixel covered twice or not at all along the edges the clipper made
    fv, fi = _ground_plane(rng, 8, 12, 6.0, 40.0, -8.0)
    fv[:, 4:] *= 0.25
    scenes.append(Scene("clip_floor_192_additive", 200, 120, fv, fi, pos_comps=4, color_offset=16,
                        state=pipeline_state(blend=util.ADDITIVE_BLEND)))
    # 40 x 30 quads = 2400 triangles (binned path), depth + alpha blending, then two instances of it, then one band of it
    fv, fi = _ground_plane(rng, 40, 30, 8.0, 60.0, -6.0)
    scenes.append(Scene("clip_floor_2400_depth_blend", 320, 200, fv, fi, pos_comps=4, color_offset=16,
                        state=pipeline_state(depth=(abi.COMPARE_LESS_OR_EQUAL, 1), blend=util.ALPHA_BLEND), use_depth=True))
    fa = fv.copy()
    fa[:, 4:] *= 0.2
    scenes.append(Scene("clip_floor_2400_instances2_additive_tiled", 320, 200, fa, fi, pos_comps=4, color_offset=16, instances=2,
                        state=pipeline_state(blend=util.ADDITIVE_BLEND), color_tiling=T))
    scenes.append(Scene("clip_floor_2400_band_96_144", 320, 200, fv, fi, pos_comps=4, color_offset=16, band=(96, 144),
                        state=pipeline_state(depth=(abi.COMPARE_LESS_OR_EQUAL, 1), blend=util.ALPHA_BLEND), use_depth=True))
    # triangles far larger than the guard band (vertices millions of pixels away) around and over the frame: every plane of
    # the clipper, polygons of up to seven vertices; few triangles (direct) and, padded with a small mesh, many (binned)
    huge = []
    for k in range(40):
        c = rng.uniform(-1.5, 1.5, 2)
        r = 10.0 ** rng.uniform(3.5, 6.0)
        ang = rng.uniform(0, 2 * np.pi) + np.array([0.0, 2.1, 4.2]) + rng.uniform(-0.4, 0.4, 3)
        col = rng.uniform(0.0, 0.2, 4)
        zz = rng.uniform(0, 1)
        w = rng.uniform(0.5, 2.0, 3)
        for j in range(3):
            x, y = c[0] + r * np.cos(ang[j]) * (1.0 if j else 1e-4), c[1] + r * np.sin(ang[j]) * (1.0 if j else 1e-4)
            huge.append([x * w[j], y * w[j], zz * w[j], w[j], *col])
    huge = np.array(huge, dtype=np.float32)
    scenes.append(Scene("clip_guard_band_huge_40", 192, 128, huge, pos_comps=4, color_offset=16,
                        state=pipeline_state(blend=util.ADDITIVE_BLEND)))
    gv, gi = util.grid_mesh(16, 12, 11)
    g4 = np.concatenate([gv[:, :3], np.ones((len(gv), 1), dtype=np.float32), gv[:, 3:]], axis=1)
    mixed_v = np.concatenate([g4, huge])
    mixed_i = np.concatenate([gi[:600], np.arange(len(huge), dtype=np.uint32) + len(g4), gi[600:]])
    scenes.append(Scene("clip_guard_band_mixed_binned", 192, 128, mixed_v, mixed_i, pos_comps=4, color_offset=16,
                        state=pipeline_state(depth=(abi.COMPARE_LESS_OR_EQUAL, 1), blend=util.ALPHA_BLEND), use_depth=True))
    # one tile with more triangles than its sorted list holds (every triangle of the draw is walked), clipped ones among them
    many = []
    for k in range(4400):
        x0, y0 = rng.uniform(-1.0, 0.2, 2)
        col = [0.01, 0.02, 0.03, 0.004]
        if k % 97 == 5:  # through the eye plane
            many += [[x0, y0, 0.5, 1.0, *col], [x0 + 0.9, y0 + 0.1, 0.5, 1.0, *col], [-(x0 + 0.2) * 0.5, -(y0 + 0.8) * 0.5, -0.25, -0.5, *col]]
        else:
            many += [[x0, y0, 0.5, 1.0, *col], [x0 + 0.9, y0 + 0.1, 0.5, 1.0, *col], [x0 + 0.2, y0 + 0.8, 0.5, 1.0, *col]]
    scenes.append(Scene("clip_overflow_4400_tris_one_tile", 40, 24, np.array(many, dtype=np.float32), pos_comps=4, color_offset=16,
                        state=pipeline_state(blend=util.ADDITIVE_BLEND)))
    # ---- points and lines (R9): two triangles over four corners in snapped window coordinates ----
    PTS, LL, LS = abi.TOPOLOGY_POINT_LIST, abi.TOPOLOGY_LINE_LIST, abi.TOPOLOGY_LINE_STRIP

    def on_grid(pv, W, H, step):  # some coordinates on pixel centres / corners / sixteenths: ties of the top-left rule
        return (np.round((pv * 0.5 + 0.5) * np.array([W, H]) * step) / step / np.array([W, H])) * 2 - 1

    pts = []
    for k in range(300):
        pxy = rng.uniform(-1.05, 1.05, 2)
        if k % 3 == 0:
            pxy = on_grid(pxy, 96, 64, rng.choice([1, 2, 16]))
        pts.append([pxy[0], pxy[1], rng.uniform(0, 1), *rng.uniform(0.1, 0.3, 4)])
    pts = np.array(pts, dtype=np.float32)
    scenes.append(Scene("points_300_additive", 96, 64, pts, state=pipeline_state(topology=PTS, blend=util.ADDITIVE_BLEND)))
    scenes.append(Scene("points_40_depth_tiled_direct", 96, 64, pts[:40], state=pipeline_state(topology=PTS, depth=(abi.COMPARE_LESS, 1)),
                        use_depth=True, color_tiling=T))
    scenes.append(Scene("points_3_fused", 96, 64, pts[6:9], state=pipeline_state(topology=PTS)))
    gone = np.array([[0.0, 0.0, 0.5, -1.0, 1, 1, 1, 1], [3.0e6, 0.0, 0.5, 1.0, 1, 1, 1, 1], [0.5, 0.5, 0.5, 0.0, 1, 1, 1, 1],
                     [np.nan, 0.0, 0.5, 1.0, 1, 1, 1, 1], [0.25, -0.25, 0.5, 2.0, 1, 0.5, 0.25, 1]], dtype=np.float32)
    scenes.append(Scene("points_behind_outside_nan", 64, 48, gone, pos_comps=4, color_offset=16, state=pipeline_state(topology=PTS)))
    segs = []
    for k in range(500):
        a = rng.uniform(-1.1, 1.1, 2)
        kind = k % 10
        d = rng.uniform(-1, 1, 2) * rng.choice([0.01, 0.1, 0.6])
        if kind == 0: d[1] = 0.0          # horizontal
        if kind == 1: d[0] = 0.0          # vertical
        if kind == 2: d[1] = d[0] * 96 / 64   # exactly diagonal in pixels: the major axis is a tie
        if kind == 3: d[:] = 0.0          # zero length
        b = a + d
        if kind in (0, 1, 2, 4):
            a, b = on_grid(a, 96, 64, 2), on_grid(b, 96, 64, 2)
        ca, cb = rng.uniform(0.05, 0.25, 4), rng.uniform(0.05, 0.25, 4)
        segs += [[a[0], a[1], rng.uniform(0, 1), *ca], [b[0], b[1], rng.uniform(0, 1), *cb]]
    segs = np.array(segs, dtype=np.float32)
    scenes.append(Scene("lines_500_additive", 96, 64, segs, state=pipeline_state(topology=LL, blend=util.ADDITIVE_BLEND)))
    scenes.append(Scene("lines_60_depth_direct_band", 96, 64, segs[:120], state=pipeline_state(topology=LL, depth=(abi.COMPARE_LESS_OR_EQUAL, 1)),
                        use_depth=True, band=(16, 48)))
    scenes.append(Scene("lines_2_fused", 96, 64, segs[20:24], state=pipeline_state(topology=LL)))
    # a spiral as an indexed (uint16) strip with perspective: w grows along it, colour and depth interpolate along the major axis
    ang = np.linspace(0, 9 * np.pi, 400)
    wv = np.linspace(1.0, 3.0, 400)
    spiral = np.stack([np.cos(ang) * ang / 30 * wv, np.sin(ang) * ang / 30 * wv, np.linspace(0.1, 0.9, 400) * wv, wv,
                       0.5 + 0.5 * np.cos(ang), 0.5 + 0.5 * np.sin(ang), np.linspace(0, 1, 400), np.full(400, 0.5)], axis=1).astype(np.float32)
    scenes.append(Scene("line_strip_spiral_u16_depth_blend_instances2", 160, 120, spiral, np.arange(400, dtype=np.uint16), pos_comps=4, color_offset=16,
                        index_type=abi.INDEX_TYPE_UINT16, instances=2, use_depth=True,
                        state=pipeline_state(topology=LS, depth=(abi.COMPARE_LESS_OR_EQUAL, 1), blend=util.ALPHA_BLEND)))
    # segments through the eye plane and far beyond the guard band (cut to w >= 2^-20 and +-2^18 px first), few and many
    far = []
    for k in range(150):
        a = [rng.uniform(-0.9, 0.9), rng.uniform(-0.9, 0.9), 0.5, 1.0]
        if k % 3 == 0: b = [rng.uniform(-2, 2), rng.uniform(-2, 2), 0.25, -rng.uniform(0.1, 2.0)]   # behind the eye
        elif k % 3 == 1: b = [rng.uniform(-1, 1) * 5.0e6, rng.uniform(-1, 1) * 5.0e6, 0.75, 1.0]     # millions of pixels away
        else: b = [rng.uniform(-0.9, 0.9), rng.uniform(-0.9, 0.9), 0.5, rng.uniform(0.5, 2.0)]
        far += [[*a, *rng.uniform(0.05, 0.2, 4)], [*b, *rng.uniform(0.05, 0.2, 4)]]
    far = np.array(far, dtype=np.float32)
    scenes.append(Scene("lines_clipped_direct", 128, 96, far[:80], pos_comps=4, color_offset=16, state=pipeline_state(topology=LL, blend=util.ADDITIVE_BLEND)))
    scenes.append(Scene("lines_clipped_binned_tiled", 128, 96, np.concatenate([far, far[::-1]]), pos_comps=4, color_offset=16, color_tiling=T,
                        state=pipeline_state(topology=LL, blend=util.ADDITIVE_BLEND)))
    # bands of draws large enough for band_prefilter_kernel (>= 65536 triangles, the band under half of the viewport's rows):
    # blocks of 256 triangles that miss the band are skipped by setup and binning.  A plain mesh; a floor through the eye plane
    # (clipped primitives inside and outside the band); a frame so small that every tile's list overflows (the raster kernel
    # then walks all triangles of the draw, skipped blocks included)
    v, i = util.grid_mesh(256, 160, 9)
    scenes.append(Scene("band_prefilter_mesh_81920", 512, 320, v, i, band=(112, 176), use_depth=True,
                        state=pipeline_state(depth=(abi.COMPARE_LESS, 1), blend=util.ALPHA_BLEND)))
    fv, fi = _ground_plane(rng, 256, 140, 8.0, 60.0, -6.0)
    scenes.append(Scene("band_prefilter_clip_floor_71680", 320, 200, fv, fi, pos_comps=4, color_offset=16, band=(128, 176), use_depth=True,
                        state=pipeline_state(depth=(abi.COMPARE_LESS_OR_EQUAL, 1), blend=util.ALPHA_BLEND)))
    va = v.copy()
    va[:, 3:] *= 0.05
    scenes.append(Scene("band_prefilter_overflow_tiles", 128, 80, va, i, band=(32, 48), state=pipeline_state(blend=util.ADDITIVE_BLEND)))
    # the bins themselves overflow (more than 8 tiles per triangle on average on a queue that has drawn nothing larger): the
    # tiles without a list walk every triangle of the draw — 300 triangles as large as the frame, some through the eye plane;
    # and 66 048 wide slivers, half of them on a band (the walk must skip the blocks the band prefilter dropped)
    big = []
    for k in range(300):
        cx, cy = rng.uniform(-0.3, 0.3, 2)
        col = [0.004, 0.008, 0.012, 0.004]
        wz = -0.5 if k % 37 == 5 else 1.0
        big += [[cx - 1.6, cy - 1.4, 0.5, 1.0, *col], [cx + 1.7, cy - 1.2, 0.5, 1.0, *col], [(cx + 0.1) * wz, (cy + 1.8) * wz, 0.5 * wz, wz, *col]]
    scenes.append(Scene("bins_overflow_300_large_clipped", 512, 256, np.array(big, dtype=np.float32), pos_comps=4, color_offset=16,
                        state=pipeline_state(blend=util.ADDITIVE_BLEND)))
    sl = np.zeros((66048 * 3, 7), dtype=np.float32)
    yy = np.repeat(np.where(np.arange(66048) % 512 < 256, rng.uniform(-0.1, 0.1, 66048), rng.uniform(0.4, 0.95, 66048)), 3)
    sl[:, 0] = np.tile([-0.98, 0.98, 0.0], 66048) + np.repeat(rng.uniform(-0.02, 0.02, 66048), 3)
    sl[:, 1] = yy + np.tile([0.0, 0.0, 0.012], 66048)
    sl[:, 2] = 0.5
    sl[:, 3:] = [0.002, 0.004, 0.006, 0.002]
    scenes.append(Scene("bins_overflow_band_prefilter_66048_slivers", 2048, 256, sl, band=(96, 160), state=pipeline_state(blend=util.ADDITIVE_BLEND)))
    return scenes
