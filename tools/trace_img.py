"""One fused image-side fprop launch at the bench size with B200GAN_IMG_DBG=16|x: CTA 0 prints its timeline."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200gan  # noqa
from b200gan import engine as E, _capi as K
from tests.parity import make_param, dev
E.begin()
g = torch.Generator().manual_seed(0)
geom = E.conv_geom(512, 32, 32, 3, 208, 5, 2)
x = dev(torch.randn(512, 32, 32, 3, generator=g))
Wp = make_param(torch.randn(5, 5, 3, 208, generator=g) * 0.05); bp = make_param(torch.randn(208, generator=g))
for i in range(2):
    y = E.conv_like("fprop", x, Wp, geom, bias=bp, act=K.ACT_LRELU, leak=0.2)
    torch.cuda.synchronize()
    print("----", flush=True)
