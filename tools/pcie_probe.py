"""Host<->device copy bandwidth probe (pinned memory), for reading the e2e number: python tools/pcie_probe.py"""
import sys; sys.path.insert(0, ".")
import time, torch, ctypes as C
from kazan_b200 import host
dev = host.Device(0); s = torch.cuda.Stream(); q = dev.wrap_stream(s.cuda_stream)
lib = host.load_library()
for nbytes in (33177600, 1 << 28):
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    h_t = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    p = C.c_void_p(); lib.kzn_host_alloc_pinned(nbytes, C.byref(p))
    for name, hp in (("torch.pin_memory", h_t.data_ptr()), ("cudaHostAlloc", p.value)):
        for direction in ("d2h", "h2d"):
            fn = (lambda: q.download(hp, d.data_ptr(), nbytes)) if direction == "d2h" else (lambda: q.upload(d.data_ptr(), hp, nbytes))
            fn(); q.wait_idle()
            t0 = time.perf_counter()
            for _ in range(5): fn()
            q.wait_idle()
            dt = (time.perf_counter() - t0) / 5
            print(f"{nbytes>>20:5d} MiB {name:17s} {direction}: {nbytes/dt/1e9:6.1f} GB/s")
    lib.kzn_host_free_pinned(p)
