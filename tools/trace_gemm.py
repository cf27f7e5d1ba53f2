"""Print the phase clock stamps of CTA (0,0,0) of the conv GEMMs at the bench shapes (B200GAN_GEMM_TRACE=1)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200gan  # noqa
from b200gan import engine as E, _capi as K
from tests.parity import make_param, dev
E.begin()
for (N, H, Cin, Cout) in [(512, 16, 208, 400), (512, 8, 400, 800)]:
    g = torch.Generator().manual_seed(0)
    geom = E.conv_geom(N, H, H, Cin, Cout, 5, 2)
    x = dev(torch.randn(N, H, H, Cin, generator=g)); dy = dev(torch.randn(N, geom.Ho, geom.Wo, Cout, generator=g))
    Wp = make_param(torch.randn(5, 5, Cin, Cout, generator=g) * 0.05); bp = make_param(torch.randn(Cout, generator=g))
    for i in range(2):
        print("== fprop", Cin, Cout, i, flush=True)
        y = E.conv_like("fprop", x, Wp, geom, bias=bp, act=K.ACT_LRELU, leak=0.2); torch.cuda.synchronize()
    for i in range(2):
        print("== dgrad", Cin, Cout, i, flush=True)
        E.conv_like("dgrad", dy, Wp, geom); torch.cuda.synchronize()

    print("== wgrad", Cin, Cout, flush=True)
    for i in range(2):
        ws, wsb = E._workspace(geom, 2)
        E.launch("b200_conv2d_wgrad", E._p(x.buf), E._p(dy.buf), E._p(Wp.g32), E.C.byref(geom), 1.0, E._p(ws), wsb, 0); torch.cuda.synchronize()
