"""Peer-copy probe on 2 GPUs: bandwidth of copy-engine pushes GPU1 -> GPU0 for a 4K band (33 MB) and for 256 MB, on one stream
and split over two streams.   gpurun --gpus 2 -- python tools/peer_probe.py"""
import torch
assert torch.cuda.device_count() >= 2
def run(nbytes, streams, reps=50):
    src = torch.empty(nbytes, dtype=torch.uint8, device="cuda:1")
    dst = torch.empty(nbytes, dtype=torch.uint8, device="cuda:0")
    ss = [torch.cuda.Stream(device="cuda:1") for _ in range(streams)]
    part = nbytes // streams
    def once():
        for k, s in enumerate(ss):
            with torch.cuda.stream(s):
                dst[k * part:(k + 1) * part].copy_(src[k * part:(k + 1) * part], non_blocking=True)
    for _ in range(5): once()
    torch.cuda.synchronize("cuda:1"); torch.cuda.synchronize("cuda:0")
    e0 = [torch.cuda.Event(enable_timing=True) for _ in ss]; e1 = [torch.cuda.Event(enable_timing=True) for _ in ss]
    with torch.cuda.device(1):
        for s, a in zip(ss, e0): a.record(s)
        for _ in range(reps): once()
        for s, b in zip(ss, e1): b.record(s)
        torch.cuda.synchronize("cuda:1")
    ms = max(a.elapsed_time(b) for a, b in zip(e0, e1)) / reps
    return nbytes / ms / 1e6
for n in (33177600, 268435456):
    for st in (1, 2, 4):
        print(f"{n/1e6:8.1f} MB  {st} stream(s): {run(n, st):7.1f} GB/s")
