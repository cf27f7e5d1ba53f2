"""Micro-benchmark: one launch at a time, L2 flushed (or not) before each, CUDA-event timed."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200gan  # noqa
from b200gan import engine as E, _capi as K
from tests.parity import make_param, dev

E.begin()
flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")

def timed(fn, flush, reps=8):
    ts = []
    for _ in range(reps):
        if flush:
            flush_buf.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]

for (N, H, Cin, Cout) in [(512, 16, 200, 400), (512, 8, 400, 800), (512, 16, 256, 512)]:
    g = torch.Generator().manual_seed(0)
    geom = E.conv_geom(N, H, H, Cin, Cout, 5, 2)
    x = dev(torch.randn(N, H, H, Cin, generator=g)); dy = dev(torch.randn(N, geom.Ho, geom.Wo, Cout, generator=g))
    Wp = make_param(torch.randn(5, 5, Cin, Cout, generator=g) * 0.05)
    fl = 2.0 * N * geom.Ho * geom.Wo * 25 * Cin * Cout
    ops = {"fprop": lambda: E.conv_like("fprop", x, Wp, geom), "dgrad": lambda: E.conv_like("dgrad", dy, Wp, geom),
           "wgrad": lambda: E.launch("b200_conv2d_wgrad", E._p(x.buf), E._p(dy.buf), E._p(Wp.g32), E.C.byref(geom), 1.0, None, 0, 0)}
    for name, fn in ops.items():
        for _ in range(3): fn()
        warm = timed(fn, False); cold = timed(fn, True)
        print("H%2d %4d->%4d %-5s warm %.3f ms (%5.0f TF/s)  L2-flushed %.3f ms (%5.0f TF/s)" % (H, Cin, Cout, name, warm, fl / warm / 1e9, cold, fl / cold / 1e9), flush=True)
