"""GPU parity: the CUDA draw path through the C ABI against the CPU oracle.

Coverage, depth results and statistics are integer work: bit exact.  Shaded colour is fp32 -> UNORM8 and the
kernels follow the oracle's operation order (R1-R8), so the bar used here is stricter than north_star's
(<= 1 LSB): byte-exact colour, and the 1 LSB tolerance is only reported if exactness ever fails.
"""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

from kazan_b200 import abi
from tests import cases, util

pytestmark = pytest.mark.gpu
GOLDEN = json.loads((Path(__file__).parent / "golden" / "oracle_golden.json").read_text())
SCENES = cases.draw_scenes()


def sha(b):
    return hashlib.sha256(b).hexdigest()


@pytest.mark.parametrize("scene", SCENES, ids=lambda s: s.name)
def test_draw_matches_oracle(gpu, scene):
    oc, od, os_ = scene.run_oracle()
    gc, gd, gs = scene.run_gpu(gpu)
    assert (gs.triangles_in, gs.triangles_culled) == (os_.triangles_in, os_.triangles_culled)
    assert (gs.triangles_clipped, gs.clip_triangles, gs.triangles_discarded, gs.clip_overflow) == \
        (os_.triangles_clipped, os_.clip_triangles, os_.triangles_discarded, 0), "R3: clipped / discarded primitives differ"
    assert (gs.fragments_covered, gs.fragments_depth_passed) == (os_.fragments_covered, os_.fragments_depth_passed), \
        "coverage / depth-test outcome differs (integer path must be bit exact)"
    assert gd.tobytes() == od.tobytes(), "depth buffer differs (must be bit exact)"
    if gc.tobytes() != oc.tobytes():
        diff = np.abs(gc.astype(np.int16) - oc.astype(np.int16))
        pytest.fail(f"colour differs: max {diff.max()} LSB on {(diff > 0).sum()} bytes (tolerance in north_star: 1 LSB; "
                    "the kernels are written to be exact)")
    assert sha(gc.tobytes()) == GOLDEN[f"draw/{scene.name}/color"]
    assert sha(gd.tobytes()) == GOLDEN[f"draw/{scene.name}/depth"]


@pytest.mark.parametrize("scene", SCENES, ids=lambda s: s.name)
def test_draw_in_a_render_pass_with_load_op_clear_matches_oracle(gpu, scene):
    """vkCmdBeginRenderPass(LOAD_OP_CLEAR) .. draw .. vkCmdEndRenderPass on attachments that hold a byte pattern, against
    the oracle's vkCmdClearAttachments + draw: (a) render area = the draw's clip rectangle — a binned draw takes the clear
    over (bricks start from the clear values, every pixel written once, no clear kernel: the launch count says so), a
    direct draw has it performed first; (b) a render area one pixel larger where the framebuffer allows — never taken over."""
    area = scene.render_area()
    if area.extent.width == 0 or area.extent.height == 0:
        pytest.skip("empty clip rectangle")
    q = gpu.queue
    n0 = q.launch_count
    scene.run_gpu(gpu)
    plain = q.launch_count - n0
    oc, od, os_ = scene.run_oracle(render_pass=area)
    n0 = q.launch_count
    gc, gd, gs = scene.run_gpu(gpu, render_pass=area)
    launches = q.launch_count - n0
    assert (gs.fragments_covered, gs.fragments_depth_passed) == (os_.fragments_covered, os_.fragments_depth_passed)
    assert gd.tobytes() == od.tobytes(), "depth differs"
    assert gc.tobytes() == oc.tobytes(), "colour differs"
    if scene.primitives * scene.instances > 256:  # binned: the draw took the clear over
        assert launches == plain, f"the clear was not fused into the draw ({launches} launches against {plain})"
    else:
        assert launches == plain + (2 if scene.use_depth else 1)
    big = abi.Rect2D(abi.Offset2D(max(0, area.offset.x - 1), area.offset.y),
                     abi.Extent2D(min(scene.width - max(0, area.offset.x - 1), area.extent.width + 2), area.extent.height))
    if big.extent.width != area.extent.width:
        oc, od, _ = scene.run_oracle(render_pass=big)
        n0 = q.launch_count
        gc, gd, _ = scene.run_gpu(gpu, render_pass=big)
        assert q.launch_count - n0 == plain + (2 if scene.use_depth else 1)
        assert gd.tobytes() == od.tobytes() and gc.tobytes() == oc.tobytes(), "larger render area"


def test_render_pass_rules(gpu):
    """A render pass without a draw still clears (at vkCmdEndRenderPass, or before the next other command); begin inside a
    render pass and end outside one are refused; LOAD leaves the attachments alone."""
    from kazan_b200 import host
    w, h = 96, 40
    q = gpu.queue
    img = gpu.buffer(w * h * 4, np.full(w * h * 4, 0xA5, dtype=np.uint8))
    dep = gpu.buffer(w * h * 4, np.full(w * h * 4, 0x11, dtype=np.uint8))
    fb = abi.Framebuffer()
    fb.color = abi.ImageDesc(img.address, abi.FORMAT_R8G8B8A8_UNORM, 0, w, h, 1, 1, 1, 0)
    fb.depth = abi.ImageDesc(dep.address, abi.FORMAT_D32_SFLOAT, 0, w, h, 1, 1, 1, 0)
    area = abi.Rect2D(abi.Offset2D(8, 4), abi.Extent2D(50, 20))
    q.cmd_begin_render_pass(fb, area, (0.25, 0.5, 0.75, 1.0), 0.5)
    with pytest.raises(host.KazanError):
        q.cmd_begin_render_pass(fb, area, (0, 0, 0, 0), None)
    q.cmd_end_render_pass()
    with pytest.raises(host.KazanError):
        q.cmd_end_render_pass()
    q.wait_idle()
    c = img.download().view(np.uint32).reshape(h, w)
    d = dep.download().view(np.float32).reshape(h, w)
    inside = np.zeros((h, w), dtype=bool)
    inside[4:24, 8:58] = True
    assert np.all(c[inside] == 0xFFBF8040) and np.all(c[~inside] == 0xA5A5A5A5)
    assert np.all(d[inside] == 0.5) and np.all(dep.download().view(np.uint32).reshape(h, w)[~inside] == 0x11111111)
    # LOAD: nothing is written; a band cuts the cleared rows
    q.cmd_begin_render_pass(fb, area, None, None)
    q.cmd_end_render_pass()
    fb.band_y0, fb.band_y1 = 16, 32
    q.cmd_begin_render_pass(fb, area, (1.0, 1.0, 1.0, 1.0), None)
    q.cmd_fill_buffer(dep.address, w * h * 4, 0, 16, 0)  # any other command performs the pending clear first
    q.wait_idle()
    c2 = img.download().view(np.uint32).reshape(h, w)
    band = np.zeros((h, w), dtype=bool)
    band[16:24, 8:58] = True
    assert np.all(c2[band] == 0xFFFFFFFF) and np.array_equal(c2[~band], c[~band])
    q.cmd_end_render_pass()
    gpu.release()


def _on_a_fresh_queue(gpu, scene):
    """The raster scratch grows with the largest draw a queue has seen; whether the bins overflow depends on it, so the
    scenes that are about overflow run on a queue of their own."""
    old = gpu.queue
    gpu.queue = gpu.device.create_queue()
    try:
        return scene.run_gpu(gpu)
    finally:
        gpu.queue.wait_idle()
        gpu.queue.destroy()
        gpu.queue = old


def test_binned_long_list_and_overflow_paths_were_exercised(gpu):
    by_name = {s.name: s for s in SCENES}
    # lists longer than the shared-memory sort holds are sorted in global memory: no tile is left without a list
    for name in ("overflow_4500_tris_one_tile", "clip_overflow_4400_tris_one_tile", "band_prefilter_overflow_tiles"):
        _, _, st = by_name[name].run_gpu(gpu)
        assert st.bin_entries > 4096 and st.bin_overflow_tiles == 0, name
    _, _, st = by_name["mesh_200x100_depth_blend"].run_gpu(gpu)
    assert st.bin_entries > 0 and st.bin_overflow_tiles == 0
    # the bins themselves overflow: tiles without a list walk every triangle of the draw (also its clipped primitives'
    # extra triangles, and not the blocks the band prefilter dropped)
    for name in ("bins_overflow_300_large_clipped", "bins_overflow_band_prefilter_66048_slivers"):
        sc = by_name[name]
        oc, od, os_ = sc.run_oracle()
        gc, gd, st = _on_a_fresh_queue(gpu, sc)
        assert st.bin_overflow_tiles > 0, f"{name}: the bins did not overflow"
        assert gc.tobytes() == oc.tobytes() and (st.fragments_covered, st.triangles_culled) == (os_.fragments_covered, os_.triangles_culled), name
    _, _, st = _on_a_fresh_queue(gpu, by_name["bins_overflow_300_large_clipped"])
    assert st.triangles_clipped > 0 and st.clip_triangles > 0


def test_band_prefilter_ran_on_the_large_band_scenes(gpu):
    """band_code_kernel + band_prefilter_kernel: two more launches than a binned draw has otherwise (setup, scan, fill, sort,
    raster), on draws of >= 65536 triangles whose band keeps under half of the viewport's rows — and only there."""
    by_name = {s.name: s for s in SCENES}
    q = gpu.queue
    for name, want in (("band_prefilter_mesh_81920", 7), ("band_prefilter_clip_floor_71680", 7), ("band_prefilter_overflow_tiles", 7),
                       ("clip_floor_2400_band_96_144", 5), ("mesh_200x100_depth_blend", 5)):
        n0 = q.launch_count
        _, _, st = by_name[name].run_gpu(gpu)
        assert q.launch_count - n0 == want, name
        if name == "band_prefilter_clip_floor_71680":
            assert st.triangles_clipped > 0


def test_clipped_primitives_took_every_path(gpu):
    """R3 on the GPU: the fused direct path (2 triangles), the direct path behind setup_kernel, the binned path with extra
    records (also under instancing and on a band), and the walk of all triangles for a tile whose list overflowed"""
    by_name = {s.name: s for s in SCENES}
    for name in ("clip_floor_quad_eye_plane", "clip_floor_192_additive", "clip_floor_2400_depth_blend", "clip_guard_band_mixed_binned"):
        _, _, st = by_name[name].run_gpu(gpu)
        assert st.triangles_clipped > 0 and st.clip_triangles > st.triangles_clipped // 4 and st.clip_overflow == 0, name
    _, _, st = by_name["clip_overflow_4400_tris_one_tile"].run_gpu(gpu)
    assert st.clip_triangles > 0


def test_full_size_fill_4k_properties(gpu):
    """BASELINE configs[2] at full size: every pixel shaded exactly once, alpha 255, rows match the oracle on a
    sample of rows (the full oracle frame is checked in bench.py's cpu_baseline leg)."""
    import ctypes as C
    from oracle import binding as ob
    W, H = 3840, 2160
    sc = util.Scene("fill4k", W, H, util.FULLSCREEN_TRIANGLE, state=util.pipeline_state(fs=abi.FS_PROCEDURAL_ORBIT))
    gc, _, st = sc.run_gpu(gpu)
    assert st.fragments_covered == W * H == st.fragments_depth_passed
    px = gc.view(np.uint32)[:W * H].reshape(H, W)
    assert np.all(px >> 24 == 255)
    lib = ob.load()
    out = np.zeros(4, dtype=np.float32)
    rng = np.random.default_rng(3)
    for _ in range(2000):
        x, y = int(rng.integers(0, W)), int(rng.integers(0, H))
        lib.kzo_fs_procedural_orbit(x, y, C.c_float(W), C.c_float(H), out.ctypes.data_as(C.POINTER(C.c_float)))
        assert int(px[y, x]) == lib.kzo_pack_rgba8(out.ctypes.data_as(C.POINTER(C.c_float))), (x, y)


def test_split_frame_bands_compose_to_the_full_frame(gpu):
    """Sort-first split: rendering N bands separately gives exactly the single-pass frame."""
    base = next(s for s in SCENES if s.name == "mesh_200x100_depth_blend")
    full_c, full_d, _ = base.run_gpu(gpu)
    H, W = base.height, base.width
    comp_c = np.zeros_like(full_c)
    comp_d = np.zeros_like(full_d)
    bands = [(0, 70), (70, 135), (135, 203), (203, 270)]
    for y0, y1 in bands:
        sc = util.Scene("band", W, H, base.vertices, base.indices, state=base.state, use_depth=True, band=(y0, y1))
        c, d, _ = sc.run_gpu(gpu)
        comp_c[y0 * W * 4:y1 * W * 4] = c[y0 * W * 4:y1 * W * 4]
        comp_d[y0 * W * 4:y1 * W * 4] = d[y0 * W * 4:y1 * W * 4]
    assert comp_c.tobytes() == full_c.tobytes() and comp_d.tobytes() == full_d.tobytes()


def test_draw_is_idempotent_for_opaque_state(gpu):
    sc = next(s for s in SCENES if s.name == "mesh_cull_back_cw")
    a = sc.run_gpu(gpu)[0].tobytes()
    b = sc.run_gpu(gpu)[0].tobytes()
    assert a == b


def test_render_area_limit_of_the_packed_pixel_boxes(gpu):
    """Pixel boxes are packed into 16 bits per coordinate: a draw whose clip rectangle reaches beyond pixel
    65535 is refused with FEATURE_NOT_PRESENT, and on the same 70000-pixel-wide image a scissor that ends at pixel 65535
    renders exactly like the oracle — including triangles that straddle the scissor edge."""
    from kazan_b200 import host
    W, H = 70000, 3
    rng = np.random.default_rng(65535)
    tris = []
    for k in range(300):  # > 256 triangles: the binned path; clustered around the right-hand scissor edge
        cx = rng.uniform(65400.0, 65600.0) / W * 2.0 - 1.0
        cy = rng.uniform(-1.0, 1.0)
        for _ in range(3):
            tris.append([cx + rng.uniform(-40, 40) / W * 2.0, cy + rng.uniform(-1.5, 1.5), rng.uniform(0, 1),
                         *rng.uniform(0, 1, 4)])
    verts = np.array(tris, dtype=np.float32)
    too_far = util.Scene("beyond_65535", W, H, verts, state=util.pipeline_state(blend=util.ADDITIVE_BLEND))
    with pytest.raises(host.KazanError) as e:
        too_far.run_gpu(gpu)
    assert e.value.code == abi.ERROR_FEATURE_NOT_PRESENT and "65535" in str(e.value)
    ok = util.Scene("scissored_to_65535", W, H, verts, state=util.pipeline_state(blend=util.ADDITIVE_BLEND),
                    scissor=abi.Rect2D(abi.Offset2D(0, 0), abi.Extent2D(65536, H)))
    oc, od, os_ = ok.run_oracle()
    gc, gd, gs = ok.run_gpu(gpu)
    assert os_.fragments_covered > 0 and gs.fragments_covered == os_.fragments_covered
    assert gc.tobytes() == oc.tobytes() and gd.tobytes() == od.tobytes()
    assert not gc.view(np.uint32)[:W * H].reshape(H, W)[:, 65536:].any()  # nothing right of the scissor


def _random_scene(rng, k):
    """A small draw with random state: the point is the combinations nobody wrote a scene for."""
    W, H = int(rng.integers(9, 300)), int(rng.integers(5, 90))
    n_tri = int(rng.choice([1, 2, 7, 40, 300, 700]))
    kind = rng.random()
    verts = []
    for _ in range(n_tri):
        c = rng.uniform(-1.2, 1.2, 2)
        size = rng.choice([0.02, 0.1, 0.5, 2.5]) if kind < 0.7 else rng.choice([0.004, 3.0])
        col = rng.uniform(-0.2, 1.3, 4)
        for _ in range(3):
            p = c + rng.uniform(-size, size, 2)
            if rng.random() < 0.3:  # land on pixel centres / pixel corners: edge and vertex ties
                p = (np.round((p * 0.5 + 0.5) * np.array([W, H]) * 2) / 2 / np.array([W, H])) * 2 - 1
            verts.append([p[0], p[1], rng.uniform(-0.1, 1.1), *(col + rng.uniform(-0.1, 0.1, 4))])
    verts = np.array(verts, dtype=np.float32)
    topology = int(rng.choice([0, 3, 4, 5, 1, 2, abi.TOPOLOGY_POINT_LIST])) if n_tri <= 40 else int(rng.choice([3, 3, 1, 2, abi.TOPOLOGY_POINT_LIST]))
    depth = None if rng.random() < 0.4 else (int(rng.integers(0, 8)), int(rng.integers(0, 2)))
    blend = None
    if rng.random() < 0.6:
        blend = (int(rng.integers(0, 10)), int(rng.integers(0, 10)), int(rng.integers(0, 5)),
                 int(rng.integers(0, 10)), int(rng.integers(0, 10)), int(rng.integers(0, 5)))
    logic = int(rng.integers(0, 16)) if rng.random() < 0.2 else None
    st = util.pipeline_state(fs=abi.FS_PROCEDURAL_ORBIT if rng.random() < 0.15 else abi.FS_VERTEX_COLOR,
                             cull=int(rng.integers(0, 4)), front=int(rng.integers(0, 2)), depth=depth, blend=blend,
                             write_mask=int(rng.choice([15, 15, 15, 7, 8, 5, 0])), logic_op=logic, topology=topology)
    scissor = None
    if rng.random() < 0.5:
        x0, y0 = int(rng.integers(0, W)), int(rng.integers(0, H))
        scissor = abi.Rect2D(abi.Offset2D(x0, y0), abi.Extent2D(int(rng.integers(1, W - x0 + 1)), int(rng.integers(1, H - y0 + 1))))
    viewport = None
    if rng.random() < 0.3:
        viewport = abi.Viewport(float(rng.integers(-8, 8)), float(rng.integers(-8, 8)), float(W + rng.integers(-6, 6)),
                                float(H + rng.integers(-4, 4)), float(rng.choice([0.0, 0.25])), float(rng.choice([1.0, 0.75])))
    band = (0, 0)
    if rng.random() < 0.2 and H > 16:
        y0 = 8 * int(rng.integers(0, H // 16))
        band = (y0, min(H, y0 + 8 * int(rng.integers(1, 4))))
    return util.Scene(f"random_{k}", W, H, verts, state=st, scissor=scissor, viewport=viewport, use_depth=depth is not None,
                      clear_color=tuple(rng.uniform(0, 1, 4)), clear_depth=float(rng.choice([1.0, 0.5, 0.0])),
                      color_tiling=abi.TILING_TILED if rng.random() < 0.25 else abi.TILING_LINEAR,
                      instances=int(rng.choice([1, 1, 1, 2])), band=band)


import os  # noqa: E402


@pytest.mark.parametrize("seed", [int(x) for x in os.environ.get("KZN_FUZZ_SEEDS", "4242").split(",")])
def test_random_scenes_agree_with_the_oracle(gpu, seed):
    """Differential test: 120 random draws (geometry sizes from sub-pixel to beyond the guard band of the screen, every
    depth / blend / logic / mask / cull / topology combination the dice find, scissor, offset viewports, tiled targets,
    bands, instances) — statistics, depth and colour bit for bit."""
    rng = np.random.default_rng(seed)
    for k in range(int(os.environ.get("KZN_FUZZ_SCENES", "120"))):
        sc = _random_scene(rng, k)
        oc, od, os_ = sc.run_oracle()
        gc, gd, gs = sc.run_gpu(gpu)
        what = (k, sc.width, sc.height, sc.primitives, sc.state.topology, sc.state.logic_op, sc.state.blend_enable,
                sc.state.depth_compare_op if sc.use_depth else None)
        assert (gs.triangles_in, gs.triangles_culled, gs.fragments_covered, gs.fragments_depth_passed) == \
               (os_.triangles_in, os_.triangles_culled, os_.fragments_covered, os_.fragments_depth_passed), what
        assert gd.tobytes() == od.tobytes(), what
        assert gc.tobytes() == oc.tobytes(), what
