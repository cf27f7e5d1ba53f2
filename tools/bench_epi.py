"""Micro-benchmark: epilogue variants of the conv GEMMs (none / bias+lrelu / fused mask), L2 flushed."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200gan  # noqa
from b200gan import engine as E, _capi as K
from tests.parity import make_param, dev

E.begin()
flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")

def timed(fn, reps=8):
    ts = []
    for _ in range(reps):
        flush_buf.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]

for (N, H, Cin, Cout) in [(512, 16, 200, 400), (512, 8, 400, 800)]:
    g = torch.Generator().manual_seed(0)
    geom = E.conv_geom(N, H, H, Cin, Cout, 5, 2)
    x = dev(torch.randn(N, H, H, Cin, generator=g)); dy = dev(torch.randn(N, geom.Ho, geom.Wo, Cout, generator=g))
    Wp = make_param(torch.randn(5, 5, Cin, Cout, generator=g) * 0.05); bp = make_param(torch.randn(Cout, generator=g))
    mx = dev(torch.randn(N, H, H, Cin, generator=g)); my = dev(torch.randn(N, geom.Ho, geom.Wo, Cout, generator=g))
    fl = 2.0 * N * geom.Ho * geom.Wo * 25 * Cin * Cout
    cases = {
        "fprop plain": lambda: E.conv_like("fprop", x, Wp, geom),
        "fprop bias+lrelu": lambda: E.conv_like("fprop", x, Wp, geom, bias=bp, act=K.ACT_LRELU, leak=0.2),
        "fprop mask": lambda: E.conv_like("fprop", x, Wp, geom, out_mask=(my, K.ACT_LRELU, 0.2)),
        "dgrad plain": lambda: E.conv_like("dgrad", dy, Wp, geom),
        "dgrad mask": lambda: E.conv_like("dgrad", dy, Wp, geom, out_mask=(mx, K.ACT_LRELU, 0.2)),
        "dgrad f32 out": lambda: E.conv_like("dgrad", dy, Wp, geom, out_f32=True),
    }
    for name, fn in cases.items():
        for _ in range(3): fn()
        t = timed(fn)
        print("H%2d %4d->%4d %-18s %.3f ms (%5.0f TF/s)" % (H, Cin, Cout, name, t, fl / t / 1e9), flush=True)
