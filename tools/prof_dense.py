import sys; sys.path.insert(0, ".")
import torch, bench
from kazan_b200 import abi, host
# more triangles per tile than the shared-memory sort holds (4096): a 2M- and a 4M-triangle mesh at 1920x1080 (4300 / 8600 per tile)
dev = host.Device(0); s = torch.cuda.Stream(); q = dev.wrap_stream(s.cuda_stream)
with torch.cuda.stream(s):
    for quads in ((1000, 1000), (2000, 1000)):
        r = bench.mesh_extra(torch, abi, host, q, s, quads=quads, size=(1920, 1080), seed=6, check_oracle=(quads == (1000, 1000)), compare_separate_clears=False)
        print(quads, {k: r[k] for k in ("triangles", "ms_per_frame", "bin_entries", "bin_overflow_tiles", "parity") if k in r})
