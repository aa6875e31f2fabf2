import sys; sys.path.insert(0,".")
import numpy as np, torch, os
from kazan_b200 import abi, host
dev=host.Device(0); s=torch.cuda.Stream(); q=dev.wrap_stream(s.cuda_stream)
for (w,h) in [(256,64),(3840,2160),(1040,48)]:
    tw,th=(w+15)//16*16,(h+15)//16*16
    lin=torch.randint(0,2**31-1,(h,w),dtype=torch.int32,device="cuda")
    til=torch.zeros(tw*th,dtype=torch.int32,device="cuda")
    back=torch.zeros(h,w,dtype=torch.int32,device="cuda")
    L=host.Image(lin.data_ptr(),w,h); T=host.Image(til.data_ptr(),w,h,tiling=abi.TILING_TILED); B=host.Image(back.data_ptr(),w,h)
    with torch.cuda.stream(s):
        n0=q.launch_count
        q.cmd_convert_tiling(L,T); q.cmd_convert_tiling(T,B)
    torch.cuda.synchronize()
    ref=lin.view(h//16,16,w//16,16).permute(0,2,1,3).contiguous().view(-1)
    print(w,h,"tiled ok",bool(torch.equal(ref,til)),"round trip ok",bool(torch.equal(back,lin)))
