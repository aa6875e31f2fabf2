import sys; sys.path.insert(0,".")
import torch, bench
from kazan_b200 import abi, host
# the 7680x4320 / 4M-triangle frame of configs[4] on one GPU (one band)
dev=host.Device(0); s=torch.cuda.Stream(); q=dev.wrap_stream(s.cuda_stream)
with torch.cuda.stream(s):
    print(bench.mesh_extra(torch, abi, host, q, s, quads=(2000, 1000), size=(7680, 4320), seed=5, check_oracle=False, compare_separate_clears=False))
