for g in 8 16 32 64; do echo "== KZN_GRID_PER_SM=$g"; KZN_GRID_PER_SM=$g timeout 300 python - <<P 2>/dev/null
import sys, json; sys.path.insert(0,".")
import torch, bench
from kazan_b200 import abi, host
dev=host.Device(0); s=torch.cuda.Stream(); q=dev.wrap_stream(s.cuda_stream)
peak,_=bench.measured_peaks()
with torch.cuda.stream(s):
    d=bench.transfer_extra(torch, abi, host, dev, q, s, peak)
for k,v in d.items():
    print(k, {a:b for a,b in v.items() if "frac" in a})
    for kk in ("copy_b2i_variants_submit","blit_and_swizzle"):
        if kk in v: print("   ", {a:b["frac_hbm"] for a,b in v[kk].items()})
P
done
