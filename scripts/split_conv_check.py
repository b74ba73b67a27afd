import sys, os
sys.path.insert(0, "/root/repo")
import torch, torch.nn.functional as F
from emojivoice_b200 import _lib
from tests.conftest import rel_l2
ctx = _lib.Context()
CASES = [(2, 192, 181, 192, 5, 2, 1), (2, 256, 181, 768, 1, 0, 1), (3, 256, 77, 768, 3, 1, 1), (2, 768, 181, 256, 3, 1, 1), (1, 256, 300, 80, 1, 0, 1), (2, 256, 130, 256, 3, 1, 1), (1, 64, 9, 160, 3, 1, 1)]
for case in CASES:
    B, Cin, T, Cout, K, pad, dil = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(B, Cin, T, generator=g); w = torch.randn(Cout, Cin, K, generator=g) / (Cin * K) ** 0.5; b = torch.randn(Cout, generator=g)
    ref = F.conv1d(x.double(), w.double(), b.double(), padding=pad, dilation=dil)
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    err = {}
    for prec in ("tf32x3", "fp32"):
        y = torch.empty(ref.shape, device="cuda")
        ctx.check(_lib.lib().ev_test_conv1d(ctx.handle, _lib.ptr(xd), _lib.ptr(wd), _lib.ptr(bd), B, Cin, T, Cout, K, 1, pad, dil, 0, _lib.PREC[prec], _lib.ptr(y), _lib.stream_ptr()), "conv")
        err[prec] = rel_l2(y.cpu(), ref)
    print(case, os.environ.get("EV_ENC_SPLIT", "f16"), {k: f"{v:.2e}" for k, v in err.items()})
