import sys; sys.path.insert(0,".")
import numpy as np, torch, bench
from kazan_b200 import abi, host
# one band (1/8 of the rows) of the 7680x4320 / 4M-triangle frame of configs[4] on one GPU: what each rank of an 8-GPU split frame runs
W, H = 7680, 4320
band = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1632, 2176)
dev=host.Device(0); s=torch.cuda.Stream(); q=dev.wrap_stream(s.cuda_stream)
verts, idx = bench.grid_mesh(2000, 1000, 5)
with torch.cuda.stream(s):
    d_verts, d_idx = torch.from_numpy(verts).cuda(), torch.from_numpy(idx.view(np.int32)).cuda()
    color = torch.zeros(W * H, dtype=torch.int32, device="cuda"); depth = torch.ones(W * H, dtype=torch.float32, device="cuda")
    fb = abi.Framebuffer(); fb.color = host.Image(color.data_ptr(), W, H).desc(); fb.depth = host.Image(depth.data_ptr(), W, H, format=abi.FORMAT_D32_SFLOAT).desc()
    fb.band_y0, fb.band_y1 = band
    st = bench.mesh_state(abi); vp = abi.Viewport(0, 0, W, H, 0, 1); sc = abi.Rect2D(abi.Offset2D(0, 0), abi.Extent2D(W, H))
    vi = abi.VertexInput(d_verts.data_ptr(), verts.nbytes, 28, 0, 3, 12)
    def frame(k):  # the band's frame as bench.py's split_frame_extra issues it: a render pass that clears the band + the draw
        q.cmd_begin_render_pass(fb, None, (0, 0, 0, 0), 1.0)
        q.cmd_draw_indexed(st, fb, vp, sc, vi, d_idx.data_ptr(), idx.nbytes, abi.INDEX_TYPE_UINT32, len(idx))
        q.cmd_end_render_pass()
    ms = bench.time_region(torch, s, frame, 10, 3) / 10
print("band", band, "ms per draw", ms)
