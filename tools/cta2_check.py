import sys, os, torch, numpy as np
sys.path.insert(0, "/root/repo")
from vltk_b200 import stages
dev = "cuda"
g = torch.Generator().manual_seed(0)
def ref_conv(x, w, sc, sh, s, p, d, relu):
    y = torch.nn.functional.conv2d(x.float().permute(0,3,1,2).double().cpu(), w.double().cpu(), None, s, p, d).permute(0,2,3,1)
    y = y * sc.double().cpu() + sh.double().cpu()
    return torch.relu(y) if relu else y
ok = True
for (n,h,w_,cin,cout,k,s,p,d) in [(8,14,14,1024,512,1,1,0,1), (5,14,14,512,512,3,1,2,2), (3,20,33,256,256,3,1,1,1), (2,14,14,2048,512,1,1,0,1), (40,14,14,512,256,1,1,0,1)]:
    x = torch.randn(n,h,w_,cin,generator=g).to(dev).bfloat16()
    wt = (torch.randn(cout,cin,k,k,generator=g)*(2.0/(cin*k*k))**0.5).bfloat16().float().to(dev)
    sc = (torch.rand(cout,generator=g)*0.5+0.75).to(dev); sh = (torch.randn(cout,generator=g)*0.1).to(dev)
    y = stages.conv2d_nhwc(x, wt, sc, sh, None, s, p, d, True, mode="bf16", tensor_cores=True).float().cpu()
    r = ref_conv(x, wt, sc, sh, s, p, d, True)
    err = ((y.double()-r).abs() - 2.0**-7*(r.abs()+1e-2)).max().item()
    print((n,h,w_,cin,cout,k), "max excess err", err, "OK" if err <= 0 else "FAIL")
    ok &= err <= 0
print("ALL OK" if ok else "SOME FAILED")
