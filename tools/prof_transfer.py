"""ncu target: a few clears / copies at 3840x2160 and 16384x16384 through the C ABI (no timing here)."""
import sys; sys.path.insert(0, ".")
import torch
from kazan_b200 import abi, host
dev = host.Device(0); s = torch.cuda.Stream(); q = dev.wrap_stream(s.cuda_stream)
with torch.cuda.stream(s):
    for w, h in ((3840, 2160), (16384, 16384)):
        n = w * h * 4
        img = torch.empty(n, dtype=torch.uint8, device="cuda"); buf = torch.empty(n, dtype=torch.uint8, device="cuda")
        buf.random_(0, 256)
        im = host.Image(img.data_ptr(), w, h)
        region = [host.buffer_image_copy(w, h)]
        sub = [host.buffer_image_copy(w - 32, h - 16, row_length=w, offset=(16, 8, 0))]
        q.cmd_clear_color_image(im, (0.25, 0.5, 0.75, 1.0))
        q.cmd_copy_buffer_to_image(buf.data_ptr(), n, im, region)
        q.cmd_copy_image_to_buffer(im, buf.data_ptr(), n, region)
        q.cmd_copy_buffer_to_image(buf.data_ptr(), n, im, sub)
        if w == 3840:  # the rows VERDICT r1 asked for: bufferOffset 4 (realigned copy), blits, tiled <-> linear
            big = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
            q.cmd_copy_buffer_to_image(big.data_ptr(), n + 64, im, [host.buffer_image_copy(w, h, buffer_offset=4)])
            img2 = torch.empty(n, dtype=torch.uint8, device="cuda")
            im2 = host.Image(img2.data_ptr(), w, h)
            def blit_region(sw, sh):
                b = abi.ImageBlit()
                b.srcSubresource = b.dstSubresource = abi.ImageSubresourceLayers(abi.ASPECT_COLOR, 0, 0, 1)
                b.srcOffsets[0], b.srcOffsets[1] = abi.Offset3D(0, 0, 0), abi.Offset3D(sw, sh, 1)
                b.dstOffsets[0], b.dstOffsets[1] = abi.Offset3D(0, 0, 0), abi.Offset3D(w, h, 1)
                return b
            for region in (blit_region(w, h), blit_region(w // 2, h // 2)):
                for filt in (0, 1):  # NEAREST, LINEAR
                    q.cmd_blit_image(im, im2, [region], filt)
            tsize, _ = host.Image(0, w, h, tiling=abi.TILING_TILED).memory_requirements()
            timg = torch.empty(tsize, dtype=torch.uint8, device="cuda")
            tim = host.Image(timg.data_ptr(), w, h, tiling=abi.TILING_TILED)
            q.cmd_convert_tiling(im, tim)
            q.cmd_convert_tiling(tim, im2)
        torch.cuda.synchronize()
    # BASELINE configs[0] as a command buffer: 64 clears and 64 copies of 1024x1024 images, merged into one launch each
    w = h = 1024
    n = w * h * 4
    imgs = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(64)]
    bufs = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(64)]
    region = [host.buffer_image_copy(w, h)]
    with q.record() as cb_clear:
        for t in imgs:
            q.cmd_clear_color_image(host.Image(t.data_ptr(), w, h), (0.25, 0.5, 0.75, 1.0))
    with q.record() as cb_copy:
        for t, b in zip(imgs, bufs):
            q.cmd_copy_buffer_to_image(b.data_ptr(), n, host.Image(t.data_ptr(), w, h), region)
    for _ in range(2):
        q.submit([cb_clear]); q.submit([cb_copy])
    torch.cuda.synchronize()
print("ok")
