import sys, torch, torch.nn.functional as F
sys.path.insert(0, '/root/repo')
from emojivoice_b200 import _lib
ctx = _lib.Context()
def rel(a,b): return float((a.double()-b).norm()/b.norm())
for case in [(2,192,181,192,5,2,1),(2,256,181,768,1,0,1),(3,256,77,768,3,1,1),(2,768,181,256,3,1,1),(1,256,300,80,1,0,1),(32,256,181,768,3,1,1),(32,768,181,256,3,1,1)]:
    B,Cin,T,Cout,K,pad,dil = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(B,Cin,T,generator=g); w = torch.randn(Cout,Cin,K,generator=g)/(Cin*K)**0.5; b = torch.randn(Cout,generator=g)
    ref = F.conv1d(x.double(), w.double(), b.double(), padding=pad, dilation=dil)
    xd,wd,bd = x.cuda(), w.cuda(), b.cuda()
    out = {}
    for prec in ("tf32x3","fp32"):
        y = torch.empty(ref.shape, device="cuda")
        ctx.check(_lib.lib().ev_test_conv1d(ctx.handle,_lib.ptr(xd),_lib.ptr(wd),_lib.ptr(bd),B,Cin,T,Cout,K,1,pad,dil,0,_lib.PREC[prec],_lib.ptr(y),_lib.stream_ptr()),"conv")
        out[prec] = rel(y.cpu(), ref)
    print(case, {k: f"{v:.2e}" for k,v in out.items()})
