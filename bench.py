#!/usr/bin/env python
"""bench.py — headline benchmark of the kazan_b200 hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" is one frame of BASELINE.json configs[2]: a fullscreen triangle rasterised and shaded with the
procedural fragment program into a 3840x2160 R8G8B8A8_UNORM target, through the C ABI
(kzn_cmd_draw -> setup_kernel + raster_kernel).  `value` is Gpixels/s shaded with everything resident in HBM;
`e2e` is the same frame through the same call with HOST buffers (vertex upload + frame read-back inside the
timed region).  `extra` carries the second half of the metric: clear / copy GB/s against the measured HBM peak.
With N > 1 (torchrun) every rank renders one 3840x2160 band of a 3840 x (2160*N) frame (sort-first split) and the
bands are composited to rank 0 over NCCL; value = all pixels of all bands / max-over-ranks time (weak scaling).

--impl reference times the reference's CPU path.  The reference (kazan-3d/kazan) cannot be built here and leaves
this path unimplemented, so — as the task's tier rules prescribe — this arm times the CPU oracle port (oracle/)
on all host threads, on the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FRAME_W, FRAME_H = 3840, 2160
# algorithmic work of the fill-rate shader per pixel: fp32-pipe instructions (an fma counts once)
FP32_INSTR_PER_PIXEL = 148
METRIC = "Gpixels/s shaded (4K fill)"
UNIT = "Gpixels/s"


def measured_peaks() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.06)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self) -> dict:
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# workload construction (no test helpers: the bench talks to the product API only)

def fill_state(abi):
    st = abi.GraphicsPipelineState()
    st.fragment_program = abi.FS_PROCEDURAL_ORBIT
    st.color_write_mask = 15
    return st


FULLSCREEN_TRIANGLE = np.array([[-1, -1, 0, 1, 1, 1, 1], [3, -1, 0, 1, 1, 1, 1], [-1, 3, 0, 1, 1, 1, 1]], dtype=np.float32)


def framebuffer(abi, address, width, height, band=(0, 0)):
    fb = abi.Framebuffer()
    fb.color = abi.ImageDesc(address, abi.FORMAT_R8G8B8A8_UNORM, abi.TILING_LINEAR, width, height, 1, 1, 1, 0)
    fb.band_y0, fb.band_y1 = band
    return fb


def time_region(torch, stream, fn, steps, warmup, barrier=None):
    """W warm-up steps, then EXACTLY `steps` steps between two events on `stream`; returns milliseconds."""
    for k in range(warmup):
        fn(k)
    torch.cuda.synchronize()
    if barrier:
        barrier()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    start.record(stream)
    for k in range(steps):
        fn(warmup + k)
    end.record(stream)
    torch.cuda.synchronize()
    if barrier:
        barrier()
    return start.elapsed_time(end)


def transfer_extra(torch, abi, host, dev, queue, stream, hbm_peak):
    """clear / copy GB/s (the second half of BASELINE's metric) at config #1's size and at HBM-resident sizes."""
    out = {}
    sizes = [("1024x1024", 1024, 1024), ("3840x2160", 3840, 2160), ("7680x4320", 7680, 4320), ("16384x16384", 16384, 16384)]
    for label, w, h in sizes:
        nbytes = w * h * 4
        # rotate over enough images that the footprint exceeds L2 (126 MB): no L2 flush kernels needed
        n_img = max(2, min(64, (768 << 20) // nbytes))
        images = [torch.empty(nbytes, dtype=torch.uint8, device="cuda") for _ in range(n_img)]
        buffers = [torch.empty(nbytes, dtype=torch.uint8, device="cuda") for _ in range(n_img)]
        for b in buffers:
            b.random_(0, 256)
        imgs = [host.Image(t.data_ptr(), w, h) for t in images]
        region = [host.buffer_image_copy(w, h)]
        reps = max(8, min(400, (24 << 30) // nbytes))
        ms = time_region(torch, stream, lambda k: queue.cmd_clear_color_image(imgs[k % n_img], (0.25, 0.5, 0.75, 1.0)), reps, 3)
        clear = nbytes * reps / ms / 1e6
        ms = time_region(torch, stream, lambda k: queue.cmd_copy_buffer_to_image(buffers[k % n_img].data_ptr(), nbytes, imgs[k % n_img], region), reps, 3)
        copy = 2 * nbytes * reps / ms / 1e6
        ms = time_region(torch, stream, lambda k: queue.cmd_copy_image_to_buffer(imgs[k % n_img], buffers[(k + 1) % n_img].data_ptr(), nbytes, region), reps, 3)
        copy_back = 2 * nbytes * reps / ms / 1e6
        # the same three commands, kernel only (CUDA-event pair around the kernel inside the library): separates the
        # kernel's bandwidth from per-command launch cost, which dominates below ~32 MB
        kc = nbytes / kernel_ms(torch, queue, lambda k: queue.cmd_clear_color_image(imgs[k % n_img], (0.25, 0.5, 0.75, 1.0)), 10) / 1e6
        kb = 2 * nbytes / kernel_ms(torch, queue, lambda k: queue.cmd_copy_buffer_to_image(buffers[k % n_img].data_ptr(), nbytes, imgs[k % n_img], region), 10) / 1e6
        # and recorded into ONE command buffer (vkBeginCommandBuffer .. vkQueueSubmit): a replay costs one host launch
        # whatever the number of commands, which is how a Vulkan application submits them
        batch = max(2, min(64, n_img))
        with queue.record() as cb_clear:
            for k in range(batch):
                queue.cmd_clear_color_image(imgs[k % n_img], (0.25, 0.5, 0.75, 1.0))
        with queue.record() as cb_copy:
            for k in range(batch):
                queue.cmd_copy_buffer_to_image(buffers[k % n_img].data_ptr(), nbytes, imgs[k % n_img], region)
        launches_per_submit = [cb_clear.launch_count, cb_copy.launch_count]  # disjoint commands merge (transfer.cu)
        sub = max(3, reps // batch)
        ms = time_region(torch, stream, lambda k: queue.submit([cb_clear]), sub, 2)
        gclear = nbytes * batch * sub / ms / 1e6
        ms = time_region(torch, stream, lambda k: queue.submit([cb_copy]), sub, 2)
        gcopy = 2 * nbytes * batch * sub / ms / 1e6
        cb_clear.destroy(); cb_copy.destroy()
        variants = {}
        if label in ("3840x2160", "7680x4320"):
            # SURVEY.md §8d stress variants of the same copy (algorithmic bytes = 2 x 4 x extent; padding is not counted),
            # recorded 8 per command buffer like the rows above: padded source pitch, sub-rectangle, bufferOffset 4
            def recorded(regions, extent_bytes, buf_bytes, src_off=0):
                big = [torch.empty(buf_bytes, dtype=torch.uint8, device="cuda") for _ in range(min(n_img, 8))]
                with queue.record() as cb:
                    for k in range(8):
                        queue.cmd_copy_buffer_to_image(big[k % len(big)].data_ptr() + src_off, buf_bytes - src_off, imgs[k % n_img], regions)
                ms_v = time_region(torch, stream, lambda k: queue.submit([cb]), max(3, sub), 2)
                n_launch = cb.launch_count
                cb.destroy()
                return {"GBps": round(2 * extent_bytes * 8 * max(3, sub) / ms_v / 1e6, 1),
                        "frac_hbm": round(2 * extent_bytes * 8 * max(3, sub) / ms_v / 1e6 / hbm_peak, 3), "launches_per_submit": n_launch}
            pitch = w + 32
            variants["padded_pitch"] = recorded([host.buffer_image_copy(w, h, row_length=pitch)], nbytes, pitch * h * 4)
            variants["sub_rect"] = recorded([host.buffer_image_copy(w - 32, h - 16, row_length=w, offset=(16, 8, 0))],
                                            (w - 32) * (h - 16) * 4, nbytes)
            variants["buffer_offset_4"] = recorded([host.buffer_image_copy(w, h, buffer_offset=4)], nbytes, nbytes + 16)
        out[label] = {"clear_submit_GBps": round(gclear, 1), "clear_submit_frac_hbm": round(gclear / hbm_peak, 3),
                      "copy_b2i_submit_GBps": round(gcopy, 1), "copy_b2i_submit_frac_hbm": round(gcopy / hbm_peak, 3),
                      "commands_per_submit": batch, "launches_per_submit": launches_per_submit,
                      "clear_kernel_GBps": round(kc, 1), "copy_b2i_kernel_GBps": round(kb, 1),
                      "clear_GBps": round(clear, 1), "clear_frac_hbm": round(clear / hbm_peak, 3),
                      "copy_b2i_GBps": round(copy, 1), "copy_b2i_frac_hbm": round(copy / hbm_peak, 3),
                      "copy_i2b_GBps": round(copy_back, 1), "copy_i2b_frac_hbm": round(copy_back / hbm_peak, 3),
                      "bytes": nbytes, "rotating_images": n_img, "reps": reps}
        if variants:
            out[label]["copy_b2i_variants_submit"] = variants
        if label in ("3840x2160", "7680x4320"):
            # the other two transfer commands north_star names: vkCmdBlitImage (algorithmic bytes: 4 written per destination
            # texel + 4 read per source texel of the blit region) and the tiled <-> linear layout swizzle (8 per texel)
            def blit_region(sw, sh, dw, dh):
                b = abi.ImageBlit()
                b.srcSubresource = b.dstSubresource = abi.ImageSubresourceLayers(abi.ASPECT_COLOR, 0, 0, 1)
                b.srcOffsets[0], b.srcOffsets[1] = abi.Offset3D(0, 0, 0), abi.Offset3D(sw, sh, 1)
                b.dstOffsets[0], b.dstOffsets[1] = abi.Offset3D(0, 0, 0), abi.Offset3D(dw, dh, 1)
                return [b]
            rows = {}
            half = [host.Image(t.data_ptr(), w // 2, h // 2) for t in buffers]  # a half-size source in the same memory
            for name, srcs, region, moved, filt in (
                    ("blit_nearest_1to1", [host.Image(t.data_ptr(), w, h) for t in buffers], blit_region(w, h, w, h), 2 * nbytes, abi.FILTER_NEAREST),
                    ("blit_linear_1to1", [host.Image(t.data_ptr(), w, h) for t in buffers], blit_region(w, h, w, h), 2 * nbytes, abi.FILTER_LINEAR),
                    ("blit_nearest_2x", half, blit_region(w // 2, h // 2, w, h), nbytes + nbytes // 4, abi.FILTER_NEAREST),
                    ("blit_linear_2x", half, blit_region(w // 2, h // 2, w, h), nbytes + nbytes // 4, abi.FILTER_LINEAR)):
                r = max(8, reps // 4)
                ms = time_region(torch, stream, lambda k: queue.cmd_blit_image(srcs[k % n_img], imgs[(k + 1) % n_img], region, filt), r, 3)
                rows[name] = {"GBps": round(moved * r / ms / 1e6, 1), "frac_hbm": round(moved * r / ms / 1e6 / hbm_peak, 3)}
            tw, th = (w + 15) // 16 * 16, (h + 15) // 16 * 16
            tiled_mem = [torch.empty(tw * th * 4, dtype=torch.uint8, device="cuda") for _ in range(n_img)]
            tiled = [host.Image(t.data_ptr(), w, h, tiling=abi.TILING_TILED) for t in tiled_mem]
            lin = [host.Image(t.data_ptr(), w, h) for t in buffers]
            r = max(8, reps // 2)
            ms = time_region(torch, stream, lambda k: queue.cmd_convert_tiling(lin[k % n_img], tiled[(k + 1) % n_img]), r, 3)
            rows["linear_to_tiled"] = {"GBps": round(2 * nbytes * r / ms / 1e6, 1), "frac_hbm": round(2 * nbytes * r / ms / 1e6 / hbm_peak, 3)}
            ms = time_region(torch, stream, lambda k: queue.cmd_convert_tiling(tiled[k % n_img], imgs[(k + 1) % n_img]), r, 3)
            rows["tiled_to_linear"] = {"GBps": round(2 * nbytes * r / ms / 1e6, 1), "frac_hbm": round(2 * nbytes * r / ms / 1e6 / hbm_peak, 3)}
            out[label]["blit_and_swizzle"] = rows
            del tiled_mem
        del images, buffers
        torch.cuda.empty_cache()
    return out


def present_extra(torch, abi, host, dev, queue, stream):
    """vkQueuePresentKHR at 4K: a rendered image into its X11 shared memory segment, for windows of
    either byte order.  The bound is the PCIe link (what e2e's frame download also rides on), not HBM."""
    W, H = 3840, 2160
    out = {"frame": [W, H], "bound": "pcie", "note": "image in HBM -> SysV segment registered with CUDA; R8G8B8A8 windows: copy engine, B8G8R8A8 windows: "
           "byte-swapping kernel storing into the mapped segment; timed from the present call to the segment holding the frame"}
    frame = torch.randint(0, 2**31 - 1, (W * H,), dtype=torch.int32, device="cuda")
    for name, order in (("b8g8r8a8_window", abi.WINDOW_B8G8R8A8), ("r8g8b8a8_window", abi.WINDOW_R8G8B8A8)):
        sc = host.Swapchain(dev, W, H, image_count=3, window_byte_order=order)
        images = [sc.image(i) for i in range(3)]
        for im in images:
            queue.cmd_copy_buffer(frame.data_ptr(), frame.numel() * 4, im.address, frame.numel() * 4, [abi.BufferCopy(0, 0, frame.numel() * 4)])
        queue.wait_idle()
        best = None
        for rep in range(6):
            idx = sc.acquire_next_image()
            t0 = time.perf_counter()
            queue.present(sc, idx)
            sc.wait_presented(idx)
            dt = time.perf_counter() - t0
            if rep >= 2:
                best = dt if best is None else min(best, dt)
        seg = sc.segment(idx)
        got = np.ctypeslib.as_array((C.c_uint32 * (W * H)).from_address(seg.mapped))
        want = frame.cpu().numpy().view(np.uint32)
        if order == abi.WINDOW_B8G8R8A8:
            want = (want & 0xFF00FF00) | ((want & 0xFF) << 16) | ((want >> 16) & 0xFF)
        out[name] = {"ms": best * 1e3, "GBps": W * H * 4 / best / 1e9, "frame_matches": bool(np.array_equal(got, want))}
        sc.destroy()
    return out


def grid_mesh(quads_x, quads_y, seed):
    """Jittered grid mesh of SURVEY.md §8d config #4 (same numbers as oracle/raster.c kzo_make_grid_mesh, restated in
    numpy so that the bench does not call the oracle for its inputs): vertex = x,y,z,r,g,b,a f32; u32 indices."""
    vx, vy = quads_x + 1, quads_y + 1
    n = vx * vy
    with np.errstate(over="ignore"):
        k = np.arange(1, 7 * n + 1, dtype=np.uint64)
        z = np.uint64(seed) + k * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = ((z >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)).reshape(n, 7)
    f = np.float32
    i = np.tile(np.arange(vx, dtype=np.float32), vy)
    j = np.repeat(np.arange(vy, dtype=np.float32), vx)
    cell_x, cell_y = f(2.0) / f(quads_x), f(2.0) / f(quads_y)
    v = np.empty((n, 7), dtype=np.float32)
    v[:, 0] = f(-1.0) + (i + (u[:, 0] - f(0.5)) * f(3.0)) * cell_x
    v[:, 1] = f(-1.0) + (j + (u[:, 1] - f(0.5)) * f(3.0)) * cell_y
    v[:, 2] = f(0.05) + f(0.9) * u[:, 2]
    v[:, 3:6] = u[:, 3:6]
    v[:, 6] = f(0.25) + f(0.75) * u[:, 6]
    qi = np.tile(np.arange(quads_x, dtype=np.uint32), quads_y)
    qj = np.repeat(np.arange(quads_y, dtype=np.uint32), quads_x)
    v00 = qj * np.uint32(vx) + qi
    v10, v01 = v00 + 1, v00 + np.uint32(vx)
    v11 = v01 + 1
    idx = np.stack([v00, v10, v11, v00, v11, v01], axis=1).reshape(-1).astype(np.uint32)
    return v, idx


def kernel_ms(torch, queue, fn, reps=30):
    """Average device duration of the dominant kernel of `fn`'s command, from the library's CUDA-event pair."""
    queue.set_kernel_timing(True)
    total = 0.0
    for k in range(reps):
        # three commands back to back, the last one read: its event pair is recorded while the GPU is still busy with
        # the earlier ones, so the pair brackets the kernel only (not the idle-stream launch latency)
        fn(3 * k)
        fn(3 * k + 1)
        fn(3 * k + 2)
        total += queue.last_kernel_ms()
    queue.set_kernel_timing(False)
    return total / reps


def roofline_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture (profiles/)."""
    f = ROOT / "profiles" / "roofline_traffic.json"
    try:
        return json.loads(f.read_text())[kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


def dispatch_extra(torch, queue, stream):
    """BASELINE configs[1]: a compute shader over 2^24 invocations through the PTX back end (IR -> PTX -> cuLaunchKernel)."""
    from kazan_b200 import backend
    out = {}
    n = 1 << 24
    golden = json.loads((ROOT / "tests" / "golden" / "ir_goldens.json").read_text())["spirv_to_ir"][0]
    words = np.frombuffer(bytes.fromhex(golden["spv_hex"]), dtype="<u4").tolist()
    module, local_size = backend.translate_spirv_to_ir(words, "main", backend.EXECUTION_MODEL_GLCOMPUTE)
    pipe = backend.ComputePipeline(queue.device, module, local_size)
    ms = time_region(torch, stream, lambda k: queue.cmd_dispatch(pipe, (65536, 256, 1)), 50, 3) / 50
    out["trivial_test_spv"] = {"invocations": n, "ms": ms, "Ginvocations_per_s": n / ms / 1e6}
    pipe.destroy()
    text = (ROOT / "tests" / "ir_programs" / "sum_loop.kazan-ir").read_text()
    pipe = backend.ComputePipeline(queue.device, backend.IrModule.parse(text))
    x = torch.randint(0, 2**31 - 1, (n,), dtype=torch.int32, device="cuda")
    y = torch.empty(n, dtype=torch.int32, device="cuda")
    ms = time_region(torch, stream, lambda k: queue.cmd_dispatch(pipe, (65536, 256, 1), x.data_ptr(), y.data_ptr()), 20, 3) / 20
    torch.cuda.synchronize()
    ok = bool(torch.equal(y, x * 256))  # int32 wrap-around, the loop's closed form
    # The IR loop is 256 iterations of two adds and a compare; ptxas folds the counter away and unrolls the accumulation into
    # 129 IADD3 (three-input adds) per invocation — profiles/sass_summary.txt, `kzn_main of sum_loop.kazan-ir` — so the rate
    # that means something is EXECUTED IADD3 per second against the integer ALU issue rate: 64 lanes per SM per clock
    # (B300_MICROARCH.md: alu pipe, one warp instruction per 2 cycles per SM sub-partition).
    iadd3 = 129
    sm_count = queue.device.sm_count
    peak = sm_count * 64 * 1.965e9 / 1e12
    rate = n * iadd3 / ms / 1e9
    out["sum_loop_256"] = {"invocations": n, "ms": ms, "Ginvocations_per_s": n / ms / 1e6,
                           "executed_iadd3_per_invocation": iadd3, "Tiadd3_per_s": rate,
                           "roofline": {"bound": "int_alu_issue", "achieved": rate, "peak": peak, "unit": "Tinstr/s", "frac": rate / peak,
                                        "peak_source": "nominal: SMs x 64 lanes x 1.965 GHz"},
                           "bytes_per_invocation": 8, "GBps": n * 8 / ms / 1e6, "parity": "closed form x*256 (mod 2^32) holds" if ok else "MISMATCH"}
    pipe.destroy()
    return out


def mesh_state(abi):
    st = abi.GraphicsPipelineState()
    st.fragment_program = abi.FS_VERTEX_COLOR
    st.color_write_mask = 15
    st.depth_test_enable = st.depth_write_enable = 1
    st.depth_compare_op = abi.COMPARE_LESS
    st.blend_enable = 1
    st.src_color_blend_factor = st.src_alpha_blend_factor = abi.BLEND_SRC_ALPHA
    st.dst_color_blend_factor = st.dst_alpha_blend_factor = abi.BLEND_ONE_MINUS_SRC_ALPHA
    return st


def mesh_extra(torch, abi, host, queue, stream, quads=(1000, 500), size=(1920, 1080), seed=4, check_oracle=True, hbm_peak=None,
               traffic_key="mesh_1M_1080p", compare_separate_clears=True):
    """BASELINE configs[3]: 1M-triangle indexed mesh, D32 LESS + alpha blend, 1920x1080.  One frame = a render pass that
    clears colour and depth (LOAD_OP_CLEAR) + vkCmdDrawIndexed."""
    W, H = size
    verts, idx = grid_mesh(*quads, seed)
    ntri = len(idx) // 3
    d_verts, d_idx = torch.from_numpy(verts).cuda(), torch.from_numpy(idx.view(np.int32)).cuda()
    color = torch.empty(W * H * 4, dtype=torch.uint8, device="cuda")
    depth = torch.empty(W * H, dtype=torch.float32, device="cuda")
    stats = torch.zeros(12, dtype=torch.int64, device="cuda")
    cimg = host.Image(color.data_ptr(), W, H)
    dimg = host.Image(depth.data_ptr(), W, H, format=abi.FORMAT_D32_SFLOAT)
    fb = abi.Framebuffer()
    fb.color, fb.depth = cimg.desc(), dimg.desc()
    st = mesh_state(abi)
    vp = abi.Viewport(0, 0, W, H, 0, 1)
    sc = abi.Rect2D(abi.Offset2D(0, 0), abi.Extent2D(W, H))
    vi = abi.VertexInput(d_verts.data_ptr(), verts.nbytes, 28, 0, 3, 12)

    def frame(k):
        # vkCmdBeginRenderPass with LOAD_OP_CLEAR on both attachments: the draw starts its bricks from the clear values and
        # writes every pixel once (no clear kernels, no framebuffer read)
        queue.cmd_begin_render_pass(fb, None, (0, 0, 0, 0), 1.0)
        queue.cmd_draw_indexed(st, fb, vp, sc, vi, d_idx.data_ptr(), idx.nbytes, abi.INDEX_TYPE_UINT32, len(idx))
        queue.cmd_end_render_pass()

    def frame_separate_clears(k):  # vkCmdClearColorImage + vkCmdClearDepthStencilImage outside a render pass, then the draw
        queue.cmd_clear_color_image(cimg, (0, 0, 0, 0))
        queue.cmd_clear_depth_stencil_image(dimg, 1.0)
        queue.cmd_draw_indexed(st, fb, vp, sc, vi, d_idx.data_ptr(), idx.nbytes, abi.INDEX_TYPE_UINT32, len(idx))

    reps = 20
    ms_separate = time_region(torch, stream, frame_separate_clears, reps, 3) / reps if compare_separate_clears else None
    color.fill_(0x5A); depth.fill_(0.25)  # the render pass must not depend on what the attachments held
    queue.set_draw_statistics(stats.data_ptr())
    frame(0)
    queue.set_draw_statistics(0)
    torch.cuda.synchronize()
    s = stats.cpu().numpy()
    ms = time_region(torch, stream, frame, reps, 3) / reps
    out = {"triangles": ntri, "frame": [W, H], "ms_per_frame": ms, "ms_per_frame_separate_clears": ms_separate, "Mtri_per_s": ntri / ms / 1e3,
           "Gpix_shaded_per_s": int(s[5]) / ms / 1e6, "fragments_covered": int(s[4]), "fragments_depth_passed": int(s[5]),
           "triangles_culled": int(s[1]), "bin_entries": int(s[3]), "bin_overflow_tiles": int(s[7]),
           "depth_complexity": int(s[4]) / (W * H)}
    # algorithmic bytes of the frame = indices + vertices + per shaded fragment (4 R + 4 W colour, 4 R + 4 W
    # depth) + per depth-failed fragment 4 R; against the measured HBM peak.  The draw is issue-bound, not bandwidth-bound:
    # the fraction says how far the frame is from the bandwidth floor, `traffic` (ncu, all kernels of the draw) how much of
    # the DRAM traffic is not algorithmic (triangle records, bins).
    alg = idx.nbytes + verts.nbytes + int(s[5]) * 16 + (int(s[4]) - int(s[5])) * 4
    if hbm_peak:
        out["roofline"] = {"bound": "hbm", "algorithmic_bytes": alg, "achieved": alg / ms / 1e6, "peak": hbm_peak, "unit": "GB/s",
                           "frac": alg / ms / 1e6 / hbm_peak, "traffic": roofline_traffic(traffic_key),
                           "note": "frame = render pass with LOAD_OP_CLEAR on colour and depth + the draw: setup + binning + sort + raster kernels (the clears happen inside raster_kernel)"}
    if check_oracle:
        from oracle import binding as ob
        lib = ob.load()
        threads = lib.kzo_hardware_threads()
        lib.kzo_set_threads(threads)
        ocolor = ob.aligned_empty(W * H * 4); ocolor[:] = 0
        odepth = ob.aligned_empty(W * H * 4); odepth.view(np.float32)[:] = 1.0
        ofb = abi.Framebuffer()
        ofb.color = abi.ImageDesc(ob.ptr(ocolor), abi.FORMAT_R8G8B8A8_UNORM, 0, W, H, 1, 1, 1, 0)
        ofb.depth = abi.ImageDesc(ob.ptr(odepth), abi.FORMAT_D32_SFLOAT, 0, W, H, 1, 1, 1, 0)
        ovi = abi.VertexInput(ob.ptr(verts), verts.nbytes, 28, 0, 3, 12)
        ostats = abi.DrawStatistics()
        t0 = time.perf_counter()
        code = lib.kzo_draw_indexed(C.byref(st), C.byref(ofb), C.byref(vp), C.byref(sc), C.byref(ovi), C.c_void_p(ob.ptr(idx)),
                                    idx.nbytes, abi.INDEX_TYPE_UINT32, len(idx), 1, 0, 0, 0, C.byref(ostats))
        dt = time.perf_counter() - t0
        assert code == 0
        same_c = bool(np.array_equal(color.cpu().numpy(), ocolor))
        same_d = bool(np.array_equal(depth.cpu().numpy().view(np.uint8), odepth))
        out["cpu_baseline"] = {"ms_per_frame": dt * 1e3, "Mtri_per_s": ntri / dt / 1e6, "cores": threads, "kind": "port"}
        out["parity"] = ("bit-exact colour+depth vs oracle at full size" if same_c and same_d else
                         f"MISMATCH colour_equal={same_c} depth_equal={same_d}")
        out["stats_match_oracle"] = (int(s[4]), int(s[5]), int(s[1])) == (ostats.fragments_covered, ostats.fragments_depth_passed,
                                                                            ostats.triangles_culled)
    return out


def oracle_fill(width, height, threads, rows=None):
    """The CPU port of the same frame (oracle/raster.c); returns (seconds, colour bytes)."""
    from oracle import binding as ob
    from kazan_b200 import abi
    lib = ob.load()
    lib.kzo_set_threads(threads)
    color = ob.aligned_empty(width * height * 4)
    color[:] = 0
    fb = framebuffer(abi, ob.ptr(color), width, height, band=(0, rows) if rows else (0, 0))
    st = fill_state(abi)
    vp = abi.Viewport(0, 0, width, height, 0, 1)
    sc = abi.Rect2D(abi.Offset2D(0, 0), abi.Extent2D(width, height))
    verts = FULLSCREEN_TRIANGLE.copy()
    vi = abi.VertexInput(ob.ptr(verts), verts.nbytes, 28, 0, 3, 12)
    t0 = time.perf_counter()
    code = lib.kzo_draw(C.byref(st), C.byref(fb), C.byref(vp), C.byref(sc), C.byref(vi), 3, 1, 0, 0, None)
    dt = time.perf_counter() - t0
    assert code == 0
    return dt, color


def run_reference(args):
    """--impl reference: the CPU path (oracle port, all host threads) on the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import binding as ob
    threads = ob.load().kzo_hardware_threads()
    # bounded sample: as many rows of the 4K frame as keep one step around a second
    probe_rows = 64
    dt, _ = oracle_fill(FRAME_W, FRAME_H, threads, rows=probe_rows)
    rows = int(min(FRAME_H, max(probe_rows, probe_rows * (1.0 / max(dt, 1e-6)))))
    rows -= rows % 8
    for k in range(args.warmup):
        oracle_fill(FRAME_W, FRAME_H, threads, rows=rows)
    t = 0.0
    for k in range(args.steps):
        dt, _ = oracle_fill(FRAME_W, FRAME_H, threads, rows=rows)
        t += dt
    value = FRAME_W * rows * args.steps / t / 1e9
    sample = f"rows [0,{rows}) of the 3840x2160 frame per step ({FRAME_W * rows} pixels)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "fullscreen-triangle procedural fragment shader at 3840x2160 R8G8B8A8_UNORM (configs[2])",
                   "note": "reference path is unimplemented!() upstream and cannot be built here; CPU oracle port timed instead"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


class FrameTargets:
    """The colour targets of a (split) frame and its per-frame composite, in the three ways the bench can run it:
    one GPU; `push` — every rank renders locally and its copy queue pushes the band into rank 0's image with the copy
    engines over NVLink peer memory while the next frame renders; `store` — the raster kernels store the band straight into
    rank 0's image; both complete a frame with flag signal/wait (kazan_b200.split_frame.PeerSplitFrame); `nccl` — the bands
    are sent to rank 0 with grouped NCCL send/recv after the draw (SplitFrame.composite)."""

    def __init__(self, torch, dist, dev, queue, width, height, world, rank, n_targets, mode):
        from kazan_b200.split_frame import PeerSplitFrame, SplitFrame
        self.mode = "single" if world == 1 else mode
        self.dist, self.world, self.rank, self.n = dist, world, rank, n_targets
        self.split = SplitFrame(width, height, world, rank)
        self.frame_bytes = width * height * 4
        self.frames = 0
        if self.mode in ("push", "store"):
            self.peer = PeerSplitFrame(dev, queue, width, height, world, rank, dist, n_targets=n_targets, mode=self.mode)
            self.addresses = self.peer.target_addresses    # the presenter's images (rank 0 reads frames from these)
            self.render = self.peer.render_addresses()     # what this rank's draws may be pointed at
        else:
            self.tensors = [torch.zeros(self.frame_bytes, dtype=torch.uint8, device="cuda") for _ in range(n_targets)]
            self.addresses = [t.data_ptr() for t in self.tensors]
            self.render = self.addresses

    def begin(self):
        """Index into `render` of the image this frame's band is drawn into."""
        if self.mode in ("push", "store"):
            return self.render.index(self.peer.begin_frame())
        return self.frames % self.n

    def end(self):
        """Publishes the band; returns the index into `addresses` of the frame's image on rank 0."""
        i = self.frames % self.n
        self.frames += 1
        if self.mode in ("push", "store"):
            self.peer.end_frame()
        elif self.mode == "nccl":
            self.split.composite(self.tensors[i], self.dist)
        return i

    def wait_idle(self, queue):
        if self.mode in ("push", "store"):
            self.peer.wait_idle()
        else:
            queue.wait_idle()

    def close(self):
        if self.mode in ("push", "store"):
            self.peer.check()
            self.peer.close()
        else:
            self.tensors = []


def split_frame_extra(torch, dist, abi, host, dev, queue, stream, world, rank, barrier, mode, hbm_peak=None):
    """BASELINE configs[4]: 7680x4320 frame of a 4M-triangle mesh (D32 LESS + alpha blend), sort-first split into `world`
    bands, composited to rank 0.  Strong scaling: the frame is fixed, every rank sees all geometry and culls to its band."""
    from kazan_b200.split_frame import SplitFrame
    W, H = 7680, 4320
    verts, idx = grid_mesh(2000, 1000, 5)
    ntri = len(idx) // 3
    d_verts, d_idx = torch.from_numpy(verts).cuda(), torch.from_numpy(idx.view(np.int32)).cuda()
    depth = torch.empty(W * H, dtype=torch.float32, device="cuda")
    ft = FrameTargets(torch, dist, dev, queue, W, H, world, rank, 2, mode)
    sf = ft.split
    y0, y1 = sf.band
    dimg = host.Image(depth.data_ptr(), W, H, format=abi.FORMAT_D32_SFLOAT)
    cimgs = [host.Image(a, W, H) for a in ft.render]
    fbs = []
    for c in cimgs:
        fb = abi.Framebuffer()
        fb.color, fb.depth = c.desc(), dimg.desc()
        fb.band_y0, fb.band_y1 = y0, y1
        fbs.append(fb)
    st = mesh_state(abi)
    vp = abi.Viewport(0, 0, W, H, 0, 1)
    sc = abi.Rect2D(abi.Offset2D(0, 0), abi.Extent2D(W, H))
    vi = abi.VertexInput(d_verts.data_ptr(), verts.nbytes, 28, 0, 3, 12)

    def frame(k):
        i = ft.begin()
        queue.cmd_begin_render_pass(fbs[i], None, (0, 0, 0, 0), 1.0)  # LOAD_OP_CLEAR on the band: the draw takes it over
        queue.cmd_draw_indexed(st, fbs[i], vp, sc, vi, d_idx.data_ptr(), idx.nbytes, abi.INDEX_TYPE_UINT32, len(idx))
        queue.cmd_end_render_pass()
        ft.end()

    reps = 20
    ms = time_region(torch, stream, frame, reps, 4, barrier) / reps
    per_rank = [ms]
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        per_rank = [float(e.item()) for e in every]
        ms = max(per_rank)
    # the band's draws alone on this rank (no composite, no other rank to wait for): what the frame would take if the exchange were free
    def band_only(k):
        queue.cmd_begin_render_pass(fbs[0], None, (0, 0, 0, 0), 1.0)
        queue.cmd_draw_indexed(st, fbs[0], vp, sc, vi, d_idx.data_ptr(), idx.nbytes, abi.INDEX_TYPE_UINT32, len(idx))
        queue.cmd_end_render_pass()
    ft.wait_idle(queue)
    ms_band = time_region(torch, stream, band_only, reps, 3, barrier) / reps
    if world > 1:
        t = torch.tensor([ms_band], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_band = float(t.item())
    out = {"frame": [W, H], "triangles": ntri, "bands": world, "ms_per_frame": ms, "ms_per_frame_by_rank": per_rank, "ms_band_draws_alone": ms_band, "Gpix_per_s": W * H / ms / 1e6, "Mtri_per_s": ntri / ms / 1e3,
            "composite_bytes": sf.composite_bytes(), "scaling": "strong", "composite": ft.mode}
    if rank == 0:  # checksum of the assembled frame (the same number whatever the band count)
        last = (ft.frames - 1) % ft.n
        h = torch.empty(W * H, dtype=torch.int32).pin_memory()
        ft.wait_idle(queue)
        queue.download(h.data_ptr(), ft.addresses[last], W * H * 4)
        queue.wait_idle()
        out["frame_checksum_rank0"] = int(h.sum(dtype=torch.int64).item())
        # full-size parity of the composited frame: the oracle renders the same 4M-triangle scene in one piece
        from oracle import binding as ob
        lib = ob.load()
        lib.kzo_set_threads(lib.kzo_hardware_threads())
        ocolor = ob.aligned_empty(W * H * 4); ocolor[:] = 0
        odepth = ob.aligned_empty(W * H * 4); odepth.view(np.float32)[:] = 1.0
        ofb = abi.Framebuffer()
        ofb.color = abi.ImageDesc(ob.ptr(ocolor), abi.FORMAT_R8G8B8A8_UNORM, 0, W, H, 1, 1, 1, 0)
        ofb.depth = abi.ImageDesc(ob.ptr(odepth), abi.FORMAT_D32_SFLOAT, 0, W, H, 1, 1, 1, 0)
        ovi = abi.VertexInput(ob.ptr(verts), verts.nbytes, 28, 0, 3, 12)
        ostats = abi.DrawStatistics()
        t0 = time.perf_counter()
        code = lib.kzo_draw_indexed(C.byref(st), C.byref(ofb), C.byref(vp), C.byref(sc), C.byref(ovi), C.c_void_p(ob.ptr(idx)),
                                    idx.nbytes, abi.INDEX_TYPE_UINT32, len(idx), 1, 0, 0, 0, C.byref(ostats))
        dt = time.perf_counter() - t0
        same = code == 0 and bool(np.array_equal(h.numpy().view(np.uint8), ocolor[:W * H * 4]))
        out["parity"] = (f"bit-exact colour vs oracle at full size ({world} band(s) composited)" if same else "MISMATCH vs oracle")
        out["cpu_baseline"] = {"ms_per_frame": dt * 1e3, "cores": int(lib.kzo_hardware_threads()), "kind": "port"}
        # SURVEY.md §8(d) bytes of the whole frame (fragment counts from the oracle's run of the same scene) against the HBM
        # peak of the GPUs that render it; geometry is replicated on every band rank, so with N bands the indices and
        # vertices are read N times (counted once here: algorithmic)
        alg = idx.nbytes + verts.nbytes + int(ostats.fragments_depth_passed) * 16 + int(ostats.fragments_covered - ostats.fragments_depth_passed) * 4
        if hbm_peak:
            out["roofline"] = {"bound": "hbm", "algorithmic_bytes": alg, "achieved": alg / ms / 1e6, "peak": hbm_peak * world, "unit": "GB/s",
                               "frac": alg / ms / 1e6 / (hbm_peak * world), "traffic": roofline_traffic("split_frame_8k") if world == 1 else None,
                               "note": "issue-bound draw: the fraction is the distance from the bandwidth floor"}
    ft.close()
    return out


def e2e_band_readback(torch, dist, abi, host, dev, queue, stream, st, vp, sc, sf, W, total_h, world, rank, barrier, steps, warm):
    """e2e for N > 1: vertex upload, draw of this rank's band, read-back of the band over this rank's PCIe link into a host
    frame shared by all ranks; two frames in flight (the read-back of frame k overlaps the draw of frame k+1)."""
    frame_bytes = W * total_h * 4
    path = f"/dev/shm/kazan_b200_e2e_{os.environ.get('MASTER_PORT', '0')}"
    if rank == 0:
        with open(path, "wb") as f:
            f.truncate(2 * frame_bytes)
    dist.barrier()
    frames = torch.from_file(path, shared=True, size=2 * frame_bytes, dtype=torch.uint8)
    a, b = sf.byte_range(rank)
    for j in range(2):  # first touch of my band's pages by my process, so that they sit in memory near my GPU's root port
        frames[j * frame_bytes + a:j * frame_bytes + b].zero_()
    dist.barrier()
    rc = torch.cuda.cudart().cudaHostRegister(frames.data_ptr(), 2 * frame_bytes, 0)
    assert int(rc) == 0, f"cudaHostRegister failed: {rc}"
    with torch.cuda.stream(stream):
        images = [dev.allocate(frame_bytes) for _ in range(2)]
        fbs = [framebuffer(abi, m.address, W, total_h, band=sf.band) for m in images]
        h_verts = torch.from_numpy(FULLSCREEN_TRIANGLE.copy()).pin_memory()
        d_verts = torch.empty(21, dtype=torch.float32, device="cuda")
        vi = abi.VertexInput(d_verts.data_ptr(), 84, 28, 0, 3, 12)
        copy_queue = dev.create_queue()
        sem_frame = [host.Semaphore(dev) for _ in range(2)]
        sem_free = [host.Semaphore(dev) for _ in range(2)]
        count = [0]

        def step(k):
            c = count[0]
            count[0] += 1
            j = c % 2
            queue.upload(d_verts.data_ptr(), h_verts.data_ptr(), 84)
            if c >= 2:
                queue.wait_semaphore(sem_free[j])  # the band drawn into this image two steps ago has been read back
            queue.cmd_draw(st, fbs[j], vp, sc, vi, 3)
            queue.signal_semaphore(sem_frame[j])
            copy_queue.wait_semaphore(sem_frame[j])
            copy_queue.download(frames.data_ptr() + j * frame_bytes + a, images[j].address + a, b - a)
            copy_queue.signal_semaphore(sem_free[j])
            if c == warm - 1 or c == warm + steps - 1:  # end of warm-up / of the timed region: the read-back is inside it
                queue.wait_semaphore(sem_free[j])
                copy_queue.signal_semaphore(sem_free[j])  # consumed above; the next ordinary wait needs it again

        ems = time_region(torch, stream, step, steps, warm, barrier)
        queue.wait_idle()
        copy_queue.wait_idle()
    t = torch.tensor([ems], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ems = float(t.item())
    dist.barrier()  # every band of the last frame is in host memory
    out = None
    if rank == 0:
        last = frames[((count[0] - 1) % 2) * frame_bytes:][:frame_bytes].numpy()
        from oracle import binding as ob
        _, want = oracle_fill(W, total_h, int(ob.load().kzo_hardware_threads()))
        same = bool(np.array_equal(last, want[:frame_bytes]))
        out = {"value": W * total_h * steps / ems / 1e6, "unit": UNIT, "h2d_bytes_per_step": 84 * world,
               "d2h_bytes_per_step": frame_bytes, "steps": steps,
               "readback": f"{world} bands, each over its own GPU's PCIe link into one shared host frame",
               "parity": "assembled host frame bit-exact vs oracle" if same else "MISMATCH vs oracle"}
    dist.barrier()
    torch.cuda.cudart().cudaHostUnregister(frames.data_ptr())
    for s_ in sem_frame + sem_free:
        s_.destroy()
    copy_queue.destroy()
    for m in images:
        m.free()
    del frames
    if rank == 0:
        try:
            os.unlink(path)
        except OSError:
            pass
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from kazan_b200 import abi, host
    from kazan_b200.split_frame import SplitFrame

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — kazan_b200 has no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    barrier = (lambda: dist.barrier()) if world > 1 else None

    dev = host.Device(local_rank)
    stream = torch.cuda.Stream()
    queue = dev.wrap_stream(stream.cuda_stream)
    hbm_peak, peak_src = measured_peaks()

    # Weak scaling: every rank shades one 3840x2160 band of a 3840 x (2160*world) frame (sort-first split); the bands are
    # composited to rank 0 in one exchange step per frame.  world == 1 is BASELINE configs[2] itself.
    W, H = FRAME_W, FRAME_H
    total_h = H * world
    frame_bytes = W * total_h * 4
    n_targets = 8  # each rank writes 8 x 33 MB of distinct lines in rotation (265 MB > 126 MiB L2): never L2 resident
    st = fill_state(abi)
    fs_program = None
    if args.fs == "compiled":
        # north_star (1)+(2): the fragment shader is a shader-compiler-ir module (the oracle's procedural shader written in
        # the IR, tests/ir_programs_v2/fs_orbit.kazan-ir) lowered to PTX by the back end and built into the raster kernels
        from kazan_b200 import backend
        ir = backend.IrModule.parse((ROOT / "tests" / "ir_programs_v2" / "fs_orbit.kazan-ir").read_text(), 2)
        fs_program = backend.FragmentProgram(dev, ir)
        ir.destroy()
        st.fragment_program = abi.FS_COMPILED
        queue.cmd_bind_fragment_program(fs_program)
        # what the shader reads: 1 / frame width, 1 / frame height (the values the built-in program derives from the framebuffer)
        queue.cmd_push_constants(np.array([np.float32(1.0) / np.float32(W), np.float32(1.0) / np.float32(H * world)], dtype=np.float32).tobytes())
    vp = abi.Viewport(0, 0, W, total_h, 0, 1)
    sc = abi.Rect2D(abi.Offset2D(0, 0), abi.Extent2D(W, total_h))
    verts = torch.from_numpy(FULLSCREEN_TRIANGLE.copy()).cuda()
    vi = abi.VertexInput(verts.data_ptr(), verts.numel() * 4, 28, 0, 3, 12)

    with torch.cuda.stream(stream):
        ft = FrameTargets(torch, dist, dev, queue, W, total_h, world, rank, n_targets, args.composite)
        sf = ft.split
        fbs = [framebuffer(abi, a, W, total_h, band=sf.band if world > 1 else (0, 0)) for a in ft.render]

        def step(k):
            i = ft.begin()
            queue.cmd_draw(st, fbs[i], vp, sc, vi, 3)
            ft.end()

        launches0 = queue.launch_count
        with ClockSampler(local_rank) as clocks:
            ms = time_region(torch, stream, step, args.steps, args.warmup, barrier)
        launches = (queue.launch_count - launches0) * args.steps // (args.steps + args.warmup)  # timed share only

    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    pixels_per_step = W * total_h
    value = pixels_per_step * args.steps / ms / 1e6  # Gpixels/s
    ms_per_step = ms / args.steps

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "fullscreen-triangle procedural fragment shader at 3840x2160 R8G8B8A8_UNORM (configs[2])"
                               + ("" if world == 1 else f"; {world} bands of 3840x2160, composite to rank 0 over NVLink peer memory ("
                                  + {"push": "copy-engine push of each band", "store": "raster kernels store into rank 0's image",
                                     "nccl": "grouped NCCL send/recv"}[args.composite] + ")"),
                   "frame": [W, total_h],
                   "fragment_program": ("procedural orbit, 32 fp32 fma iterations; " +
                                        ("compiled from shader-compiler-ir (fs_orbit.kazan-ir) by the PTX back end into the raster kernel"
                                         if fs_program else "built-in CUDA form (--fs builtin)")),
                   "l2": f"rotating {n_targets} render targets ({n_targets * W * H * 4 >> 20} MiB written per rank > 126 MiB L2)"},
        "gpu_launches": int(launches),
    }
    if world > 1:
        line["config"]["composite_bytes_per_step"] = sf.composite_bytes()
        line["config"]["composite"] = {
            "push": "bands rendered locally, pushed into rank 0's image by the copy engines over NVLink peer memory while the next frame renders; flag signal/wait per frame",
            "store": "bands stored by the raster kernels straight into rank 0's image over NVLink peer memory; flag signal/wait per frame",
            "nccl": "grouped NCCL send/recv of the bands to rank 0 after the draw"}[ft.mode]
        # the presenter ingests (world-1) bands per frame over its NVLink port: measured peer-copy bandwidth 770 GB/s per direction
        ingest_gbps = sf.composite_bytes() / (ms_per_step * 1e-3) / 1e9
        line["nvlink"] = {"bound": "presenter_ingest", "presenter_ingest_GBps": ingest_gbps, "peak": 770.0, "frac": ingest_gbps / 770.0,
                          "peak_source": "B200_PROFILING.md measured peer copy, per direction",
                          "note": "weak-scaled fill is link-bound by construction: one GPU shades 4 B/pixel x ~200 Gpixel/s = 800 GB/s of pixels, more than "
                                  "one NVLink port takes (770 GB/s measured), so composite-to-one-GPU cannot scale past one band's worth of ingest; "
                                  "extra.split_frame_8k_4Mtri (strong scaling of a geometry-bound frame) is the scaling curve to track"}

    # ---- e2e: the same frame through the same calls with HOST buffers: vertex upload, draw, composite, frame read-back.
    # Every step uploads its vertices from pinned memory and delivers its finished frame to pinned host memory; the
    # read-back runs on rank 0's copy queue (semaphores order it after the frame and free the target for reuse), so the
    # 33 MB PCIe transfer of frame k overlaps the rendering of frame k+1 — all of it inside the timed region: the last
    # step waits for the last read-back.
    with torch.cuda.stream(stream):
        h_verts = torch.from_numpy(FULLSCREEN_TRIANGLE.copy()).pin_memory()
        h_frames = [torch.empty(frame_bytes if rank == 0 else 4, dtype=torch.uint8).pin_memory() for _ in range(2)]
        d_verts = torch.empty(21, dtype=torch.float32, device="cuda")
        vi2 = abi.VertexInput(d_verts.data_ptr(), 84, 28, 0, 3, 12)
        e_steps, e_warm = max(5, min(args.steps, 20)), 3
        copy_queue = dev.create_queue() if rank == 0 else None
        sem_frame = [host.Semaphore(dev) for _ in range(n_targets)] if rank == 0 else []
        sem_free = [host.Semaphore(dev) for _ in range(n_targets)] if rank == 0 else []
        e_count = [0]

        def e2e_step(k):
            c = e_count[0]
            e_count[0] += 1
            queue.upload(d_verts.data_ptr(), h_verts.data_ptr(), 84)
            j = ft.frames % n_targets  # the image of this frame on rank 0
            if rank == 0 and c >= n_targets:
                queue.wait_semaphore(sem_free[j])  # its previous frame has been read back
            i = ft.begin()
            queue.cmd_draw(st, fbs[i], vp, sc, vi2, 3)
            ft.end()
            if rank == 0:
                queue.signal_semaphore(sem_frame[j])
                copy_queue.wait_semaphore(sem_frame[j])
                copy_queue.download(h_frames[c % 2].data_ptr(), ft.addresses[j], frame_bytes)
                copy_queue.signal_semaphore(sem_free[j])
                if c == e_warm + e_steps - 1 or c == e_warm - 1:  # end of warm-up / end of the timed region
                    queue.wait_semaphore(sem_free[j])
                    sem_free[j].resignal = True

        def e2e_step_checked(k):
            # a semaphore consumed by the end-of-region wait must be signalled again before its next ordinary wait
            e2e_step(k)
            if rank == 0:
                for sm in sem_free:
                    if getattr(sm, "resignal", False):
                        copy_queue.signal_semaphore(sm)
                        sm.resignal = False

        ems = time_region(torch, stream, e2e_step_checked, e_steps, e_warm, barrier)
        ft.wait_idle(queue)
        if copy_queue:
            copy_queue.wait_idle()
        h_frame = h_frames[(e_count[0] - 1) % 2]
    if world > 1:
        t = torch.tensor([ems], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ems = float(t.item())
    line["e2e"] = {"value": pixels_per_step * e_steps / ems / 1e6, "unit": UNIT, "h2d_bytes_per_step": 84 * world,
                   "d2h_bytes_per_step": frame_bytes, "steps": e_steps,
                   "note": "bounded by the frame read-back: 4 B/pixel over PCIe (55 GB/s measured, tools/pcie_probe.py) = 13.8 Gpixels/s"}
    if world > 1:
        # With several GPUs a frame that is wanted in HOST memory does not have to meet on one GPU first: every rank reads
        # its own band back over its own PCIe link into one host frame (a /dev/shm mapping registered with CUDA by every
        # rank).  Same calls (upload, draw, download), same timed region rules; the NVLink composite above is kept as
        # `e2e_composited` (frame assembled on the presenting GPU, then read back over that GPU's link alone).
        line["e2e_composited"] = line["e2e"]
        band_e2e = e2e_band_readback(torch, dist, abi, host, dev, queue, stream, st, vp, sc, sf, W, total_h, world, rank, barrier,
                                     e_steps, e_warm)
        if rank == 0:
            line["e2e"] = band_e2e

    if rank == 0 and world == 1:
        # ---- roofline of the dominant kernel (raster_direct_kernel), its duration measured live with the library's
        # CUDA-event pair around that kernel on the launching stream ----
        with torch.cuda.stream(stream):
            k_ms_events = kernel_ms(torch, queue, lambda k: queue.cmd_draw(st, fbs[k % n_targets], vp, sc, vi, 3))
        # The draw's kernels are launched with programmatic dependent launch: in the timed region the raster kernel's
        # launch latency is hidden behind setup_kernel, while an event pair recorded around it sits between the two kernels
        # and brackets that latency too.  The kernel cannot take longer than the whole step (setup_kernel + raster kernel,
        # back to back on one stream), so the smaller of the two is its duration (a lower bound on its rate when the step
        # time is used).
        k_ms = min(k_ms_events, ms_per_step)
        sec = k_ms / 1e3
        clk = clocks.summary()
        line["clocks"] = clk
        # The fragment program is fp32-issue bound by construction (min(HBM write, FP32 issue); 148
        # fp32 instructions against 4 bytes per pixel), so THE roofline of this kernel is the fp32 pipe.  Its peak is
        # measured in this run (kzn_calibrate_fp32_issue: packed-FMA chains only), the nominal figure is given beside it.
        with torch.cuda.stream(stream):
            measured_issue = queue.calibrate_fp32_issue()
        sm_mhz = clk["sm_mhz"] or 1965.0
        nominal_issue = dev.sm_count * 128 * sm_mhz * 1e6 / 1e12  # T fp32 lane-instructions/s at the sampled clock
        issue = W * H * FP32_INSTR_PER_PIXEL / sec / 1e12
        line["roofline"] = {"bound": "fp32_issue", "achieved": issue, "peak": nominal_issue, "unit": "Tinstr/s", "frac": issue / nominal_issue,
                            "calibrated_peak": measured_issue, "frac_of_calibrated": issue / measured_issue,
                            "traffic": roofline_traffic("raster_direct_kernel"),
                            "kernel": "kzn_fs_raster_direct_0 (raster_direct_kernel with the compiled shader)" if fs_program else "raster_direct_kernel<ORBIT,0>",
                            "kernel_ms": k_ms, "kernel_ms_event_pair": k_ms_events,
                            "kernel_ms_source": ("event pair around the kernel" if k_ms_events <= ms_per_step else
                                                 "step time: upper bound of the kernel's duration (the event pair also brackets the launch latency that programmatic dependent launch hides)"),
                            "kernel_share_of_step": k_ms / ms_per_step,
                            "fp32_instr_per_pixel": FP32_INSTR_PER_PIXEL,
                            "peak_source": "nominal: SMs x 128 fp32 lanes x the SM clock sampled during the timed region; calibrated_peak = kzn_calibrate_fp32_issue (FFMA2 chains) in this run",
                            "nominal_peak": nominal_issue, "frac_of_nominal": issue / nominal_issue, "sm_mhz": sm_mhz,
                            "note": "north_star: 'FP32 issue rate for shader-bound work'; the HBM side of the same kernel is `roofline_hbm`"}
        achieved = W * H * 4 / sec / 1e9
        line["roofline_hbm"] = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                                "traffic": roofline_traffic("raster_direct_kernel"), "peak_source": peak_src,
                                "note": "4 B written per pixel (33 177 600 B per launch); not the binding resource of this shader"}
        if args.profile:
            emit(line)
            return 0

        # ---- CPU baseline (oracle port) + full-frame parity ----
        from oracle import binding as ob
        threads = ob.load().kzo_hardware_threads()
        dt, ocolor = oracle_fill(W, H, threads)
        same = bool(np.array_equal(h_frame.numpy(), ocolor))
        line["cpu_baseline"] = {"value": W * H / dt / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "one full 3840x2160 frame, oracle/raster.c, all host threads"}
        line["parity"] = "bit-exact vs oracle on the full 4K frame" if same else "MISMATCH vs oracle"

        # ---- the same frames with the built-in CUDA form of the shader (what the compiled form is measured against) ----
        if fs_program:
            with torch.cuda.stream(stream):
                st_b = fill_state(abi)
                n_b = max(50, min(args.steps, 500))
                ms_b = time_region(torch, stream, lambda k: queue.cmd_draw(st_b, fbs[k % n_targets], vp, sc, vi, 3), n_b, 5)
            line["builtin_fragment_program"] = {"value": pixels_per_step * n_b / ms_b / 1e6, "unit": UNIT, "steps": n_b,
                                                "compiled_over_builtin": value / (pixels_per_step * n_b / ms_b / 1e6)}

        # ---- the rest of BASELINE's metric: clear / copy GB/s vs HBM peak, compute dispatch, 1M-triangle mesh ----
        with torch.cuda.stream(stream):
            line["extra"] = {"transfer": transfer_extra(torch, abi, host, dev, queue, stream, hbm_peak),
                             "hbm_peak_GBps": hbm_peak, "hbm_peak_source": peak_src,
                             "dispatch_2p24": dispatch_extra(torch, queue, stream),
                             "mesh_1M_1080p": mesh_extra(torch, abi, host, queue, stream, hbm_peak=hbm_peak),
                             "present_4k": present_extra(torch, abi, host, dev, queue, stream)}
    if not args.profile:
        with torch.cuda.stream(stream):
            ft.close()
            del fbs
            torch.cuda.empty_cache()
            sfx = split_frame_extra(torch, dist, abi, host, dev, queue, stream, world, rank, barrier, args.composite, hbm_peak)
        if rank == 0:
            line.setdefault("extra", {})["split_frame_8k_4Mtri"] = sfx
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


_RESULT_FD = None


def emit(line: dict) -> None:
    """The one JSON line of the contract, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    # stdout carries exactly one JSON line: whatever libraries print while the bench runs (NCCL's version banner when
    # NCCL_DEBUG is set, ...) goes to stderr instead
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--composite", default="push", choices=["push", "store", "nccl"],
                    help="N > 1: how bands reach rank 0 (see FrameTargets)")
    ap.add_argument("--profile", action="store_true", help="timed region only (for ncu): no e2e, CPU baseline or transfer sweep")
    ap.add_argument("--fs", default="compiled", choices=["compiled", "builtin"],
                    help="the fill-rate fragment shader: compiled from its shader-compiler-ir module by the PTX back end (default), or the built-in CUDA form")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
