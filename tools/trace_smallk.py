import os, sys, ctypes as C, torch
sys.path.insert(0, "/root/repo")
import b200gan
from b200gan import engine as E, _capi as K
from tests.parity import make_param, dev
E.begin()
N, H, Cin, Cout, k = 512, 32, 3, 200, 5
g = torch.Generator().manual_seed(0)
geom = E.conv_geom(N, H, H, Cin, Cout, k, 2)
x = dev(torch.randn(N, H, H, Cin, generator=g)); dy = dev(torch.randn(N, geom.Ho, geom.Wo, Cout, generator=g))
Wp = make_param(torch.randn(k, k, Cin, Cout, generator=g) * 0.05); bp = make_param(torch.randn(Cout, generator=g))
for i in range(2):
    print("== dgrad", i, flush=True)
    E.conv_like("dgrad", dy, Wp, geom); torch.cuda.synchronize()
for i in range(2):
    print("== fprop", i, flush=True)
    E.conv_like("fprop", x, Wp, geom, bias=bp, act=K.ACT_LRELU, leak=0.2); torch.cuda.synchronize()
