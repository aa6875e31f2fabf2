import sys; sys.path.insert(0, ".")
import ctypes as C
import numpy as np, torch, bench
from kazan_b200 import abi, host
from oracle import binding as ob
# 1M points / 1M short segments / a 1M-vertex strip at 1920x1080, additive blend: time per draw and full-size parity vs the oracle
W, H, N = 1920, 1080, 1_000_000
rng = np.random.default_rng(11)
dev = host.Device(0); s = torch.cuda.Stream(); q = dev.wrap_stream(s.cuda_stream)
st = abi.GraphicsPipelineState(); st.fragment_program = abi.FS_VERTEX_COLOR; st.color_write_mask = 15
st.blend_enable = 1; st.src_color_blend_factor = st.dst_color_blend_factor = st.src_alpha_blend_factor = st.dst_alpha_blend_factor = 1
vp = abi.Viewport(0, 0, W, H, 0, 1); sc = abi.Rect2D(abi.Offset2D(0, 0), abi.Extent2D(W, H))
a = rng.uniform(-1, 1, (N, 2)); d = rng.uniform(-1, 1, (N, 2)) * 0.01
col = rng.uniform(0.02, 0.1, (N, 4))
points = np.concatenate([a, np.full((N, 1), 0.5), col], axis=1).astype(np.float32)
segs = np.empty((2 * N, 7), dtype=np.float32)
segs[0::2] = points; segs[1::2] = np.concatenate([a + d, np.full((N, 1), 0.5), col[::-1]], axis=1)
walk = np.cumsum(rng.uniform(-1, 1, (N, 2)) * 0.004, axis=0); walk = (walk - walk.min(0)) / (walk.max(0) - walk.min(0)) * 1.9 - 0.95
strip = np.concatenate([walk, np.full((N, 1), 0.5), col], axis=1).astype(np.float32)
with torch.cuda.stream(s):
    for name, verts, topo in (("points", points, abi.TOPOLOGY_POINT_LIST), ("segments", segs, abi.TOPOLOGY_LINE_LIST), ("strip", strip, abi.TOPOLOGY_LINE_STRIP)):
        st.topology = topo
        dv = torch.from_numpy(verts).cuda()
        color = torch.zeros(W * H, dtype=torch.int32, device="cuda")
        fb = abi.Framebuffer(); fb.color = host.Image(color.data_ptr(), W, H).desc()
        vi = abi.VertexInput(dv.data_ptr(), verts.nbytes, 28, 0, 3, 12)
        def frame(k):
            q.cmd_begin_render_pass(fb, None, (0, 0, 0, 0), None)
            q.cmd_draw(st, fb, vp, sc, vi, len(verts))
            q.cmd_end_render_pass()
        ms = bench.time_region(torch, s, frame, 10, 3) / 10
        lib = ob.load(); lib.kzo_set_threads(lib.kzo_hardware_threads())
        oc = ob.aligned_empty(W * H * 4); oc[:] = 0
        ofb = abi.Framebuffer(); ofb.color = abi.ImageDesc(ob.ptr(oc), abi.FORMAT_R8G8B8A8_UNORM, 0, W, H, 1, 1, 1, 0)
        ovi = abi.VertexInput(ob.ptr(verts), verts.nbytes, 28, 0, 3, 12)
        assert lib.kzo_draw(C.byref(st), C.byref(ofb), C.byref(vp), C.byref(sc), C.byref(ovi), len(verts), 1, 0, 0, None) == 0
        same = bool(np.array_equal(color.cpu().numpy().view(np.uint8), oc))
        print(f"{name}: {len(verts)} vertices, {ms:.3f} ms per draw, {'bit-exact vs oracle' if same else 'MISMATCH'}")
