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
