"""wgrad stream-K cut: fewest 64-pixel chunks per CTA (b200_set_tuning("wgrad_min_chunks")) on the small layers of
pix2pix / VAE / cnn.  10 launches per CUDA graph, median of 5."""
import os, sys
import ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["TUNE_SETS"] = ""
import b200gan  # noqa
from b200gan import engine as E, _capi as K
from tests.parity import make_param, dev
from tools.tune_layers import timed, SETS

E.begin()
for name in ("p2p", "vae", "cnn"):
    for (N, H, Cin, Cout, k) in SETS[name]:
        g0 = torch.Generator().manual_seed(0)
        geom = E.conv_geom(N, H, H, Cin, Cout, k, 2)
        x = dev(torch.randn(N, H, H, Cin, generator=g0)); dy = dev(torch.randn(N, geom.Ho, geom.Wo, Cout, generator=g0))
        Wp = make_param(torch.randn(k, k, Cin, Cout, generator=g0) * 0.05)

        def wg():
            ws, wsb = E._workspace(geom, 2)
            E.launch("b200_conv2d_wgrad", E._p(x.buf), E._p(dy.buf), E._p(Wp.g32), C.byref(geom), 1.0, E._p(ws), wsb, 0)
        res = []
        for mc in (1, 2, 4, 8, 16, 32):
            K.set_tuning("wgrad_min_chunks", mc)
            res.append("%d:%.1f" % (mc, timed(wg)))
        K.set_tuning("wgrad_min_chunks", 0)
        print("%-4s N%d %dx%dx%d->%d k%d wgrad  chunks/CTA:us  %s" % (name, N, H, H, Cin, Cout, k, "  ".join(res)), flush=True)
