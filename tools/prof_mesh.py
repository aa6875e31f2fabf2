import sys; sys.path.insert(0,".")
import torch, bench
from kazan_b200 import abi, host
dev=host.Device(0); s=torch.cuda.Stream(); q=dev.wrap_stream(s.cuda_stream)
with torch.cuda.stream(s):
    print(bench.mesh_extra(torch, abi, host, q, s, check_oracle=False, compare_separate_clears=False))
