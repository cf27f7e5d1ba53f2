"""pix2pix PatchGAN head (512 -> 1 channel, k4 s2; hem/models/pix2pix.py:256): fprop / dgrad / wgrad timings,
10 launches per CUDA graph."""
import os, sys
import ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200gan  # noqa
from b200gan import engine as E, _capi as K
from tests.parity import make_param, dev
from tools.tune_layers import timed  # noqa  (TUNE_SETS= keeps its sweep empty)

E.begin()
N, H, Cin, Cout, k = 16, 16, 512, 1, 4
g0 = torch.Generator().manual_seed(0)
geom = E.conv_geom(N, H, H, Cin, Cout, k, 2)
x = dev(torch.randn(N, H, H, Cin, generator=g0)); dy = dev(torch.randn(N, geom.Ho, geom.Wo, Cout, generator=g0))
Wp = make_param(torch.randn(k, k, Cin, Cout, generator=g0) * 0.05)
bp = make_param(torch.randn(Cout, generator=g0))
print("routes", [K.route(geom, i) for i in range(3)])
print("fprop %.1f us" % timed(lambda: E.conv_like("fprop", x, Wp, geom, bias=bp)))
print("dgrad %.1f us" % timed(lambda: E.conv_like("dgrad", dy, Wp, geom)))
def wg():
    ws, wsb = E._workspace(geom, 2)
    E.launch("b200_conv2d_wgrad", E._p(x.buf), E._p(dy.buf), E._p(Wp.g32), C.byref(geom), 1.0, E._p(ws), wsb, 0)
print("wgrad %.1f us" % timed(wg))
