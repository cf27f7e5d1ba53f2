"""Micro-benchmark: tcgen05 conv fprop/dgrad/wgrad throughput vs channel count (TMA row alignment)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200gan  # noqa
from b200gan import engine as E, _capi as K
from tests.parity import make_param, dev

def run(N, H, W, Cin, Cout, k=5, s=2, reps=10):
    g = torch.Generator().manual_seed(0)
    geom = E.conv_geom(N, H, W, Cin, Cout, k, s)
    x = dev(torch.randn(N, H, W, Cin, generator=g)); dy = dev(torch.randn(N, geom.Ho, geom.Wo, Cout, generator=g))
    Wp = make_param(torch.randn(k, k, Cin, Cout, generator=g) * 0.05)
    fl = 2.0 * N * geom.Ho * geom.Wo * k * k * Cin * Cout
    out = {}
    for name, fn in (("fprop", lambda: E.conv_like("fprop", x, Wp, geom)), ("dgrad", lambda: E.conv_like("dgrad", dy, Wp, geom)),
                     ("wgrad", lambda: E.launch("b200_conv2d_wgrad", E._p(x.buf), E._p(dy.buf), E._p(Wp.g32), E.C.byref(geom), 1.0, None, 0, 0))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out[name] = (ms, fl / ms / 1e9)
    return out

E.begin()
for (H, Cin, Cout) in [(16, 192, 384), (16, 200, 400), (16, 256, 512), (16, 256, 416), (8, 384, 768), (8, 400, 800), (8, 448, 832), (8, 512, 1024), (8, 512, 832)]:
    r = run(512, H, H, Cin, Cout)
    print("H%2d Cin %4d Cout %4d | " % (H, Cin, Cout) + " | ".join("%s %.3f ms %6.0f TF/s" % (k_, v[0], v[1]) for k_, v in r.items()), flush=True)
