"""Launch the heavy conv-family kernels once each at the bench sizes (for ncu --set full)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200gan  # noqa
from b200gan import engine as E, _capi as K
from tests.parity import make_param, dev

# physical channel counts of the bench step (the layer API stores 200-channel layers as 208)
CASES = [(512, 16, 16, 208, 400, 5, 2), (512, 8, 8, 400, 800, 5, 2), (512, 32, 32, 3, 208, 5, 2),
         (16, 128, 128, 64, 128, 4, 2)]          # + pix2pix e2 / d7: the persistent double-buffered 2-CTA kernel
reps = int(os.environ.get("REPS", "1"))
E.begin()
for (N, H, W, Cin, Cout, k, s) in CASES:
    g = torch.Generator().manual_seed(0)
    geom = E.conv_geom(N, H, W, Cin, Cout, k, s)
    x = dev(torch.randn(N, H, W, Cin, generator=g))
    dy = dev(torch.randn(N, geom.Ho, geom.Wo, Cout, generator=g))
    Wp = make_param(torch.randn(k, k, Cin, Cout, generator=g) * 0.05)
    bp = make_param(torch.randn(Cout, generator=g))
    for _ in range(reps):
        y = E.conv_like("fprop", x, Wp, geom, bias=bp, act=K.ACT_LRELU, leak=0.2)
        gx = E.conv_like("dgrad", dy, Wp, geom)
        ws, wsb = E._workspace(geom, 2)
        E.launch("b200_conv2d_wgrad", E._p(x.buf), E._p(dy.buf), E._p(Wp.g32), E.C.byref(geom), 1.0, E._p(ws), wsb, 0)
    torch.cuda.synchronize()
print("done")
