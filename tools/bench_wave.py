"""Micro-benchmark: separate the per-CTA fixed cost from the per-tap cost of the tap GEMM by varying the
number of CTAs (batch) and of taps (filter size) at fixed tile shape.  L2 flushed between launches."""
import os, sys
import ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200gan  # noqa
from b200gan import engine as E, _capi as K
from tests.parity import make_param, dev

E.begin()
flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")

def timed(fn, reps=5, inner=10):
    """`inner` back-to-back launches captured in a CUDA graph (no host time between them); median of `reps`."""
    st = torch.cuda.Stream()
    E.S.stream = C.c_void_p(st.cuda_stream)
    keep = []
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        fn(); torch.cuda.synchronize()
        with torch.cuda.graph(graph, stream=st):
            for _ in range(inner):
                keep.append(fn())
        # cold variant: an L2 flush (512 MB memset) before every launch, all inside one graph
        cold = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cold, stream=st):
            for _ in range(inner):
                flush_buf.zero_()
                keep.append(fn())
        fl = torch.cuda.CUDAGraph()
        with torch.cuda.graph(fl, stream=st):
            for _ in range(inner):
                flush_buf.zero_()
    def med(g):
        ts = []
        for _ in range(reps):
            flush_buf.zero_(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / inner)
        ts.sort()
        return ts[len(ts) // 2]
    return med(graph), med(cold) - med(fl)

CASES = [
    # N, H, Cin, Cout, k
    (296, 16, 200, 400, 5), (592, 16, 200, 400, 5), (512, 16, 200, 400, 5),
    (296, 16, 200, 400, 3), (296, 16, 200, 400, 1),
    (296, 16, 256, 400, 5), (296, 16, 192, 400, 5), (296, 16, 256, 416, 5),
    (512, 8, 400, 800, 5), (512, 8, 400, 800, 3), (512, 8, 400, 800, 1), (512, 8, 384, 832, 5),
]
if os.environ.get("WAVE_CASES") == "step":
    CASES = [(512, 16, 200, 400, 5), (512, 8, 400, 800, 5)]
if os.environ.get("WAVE_CASES") == "ab":
    CASES = [(512, 16, 208, 400, 5), (592, 16, 208, 400, 5), (512, 8, 400, 800, 5), (592, 8, 400, 800, 5)]
if os.environ.get("WAVE_CASES") == "cin":
    CASES = [(296, 16, c, 400, 5) for c in (192, 200, 208, 224, 256, 264, 320, 328, 336)]
if os.environ.get("WAVE_CASES") == "p2p":
    # pix2pix encoder / decoder geometries at the BASELINE batch (hem/models/pix2pix.py:187-227), k4 s2
    CASES = [(16, 128, 64, 128, 4), (16, 64, 128, 256, 4), (16, 32, 256, 512, 4), (16, 16, 512, 512, 4),
             (16, 8, 512, 512, 4), (16, 32, 256, 1024, 4), (16, 16, 512, 1024, 4)]
for (N, H, Cin, Cout, k) in CASES:
    g = torch.Generator().manual_seed(0)
    geom = E.conv_geom(N, H, H, Cin, Cout, k, 2)
    x = dev(torch.randn(N, H, H, Cin, generator=g)); dy = dev(torch.randn(N, geom.Ho, geom.Wo, Cout, generator=g))
    Wp = make_param(torch.randn(k, k, Cin, Cout, generator=g) * 0.05)
    fl = 2.0 * N * geom.Ho * geom.Wo * k * k * Cin * Cout
    variants = [("fprop", lambda: E.conv_like("fprop", x, Wp, geom)), ("dgrad", lambda: E.conv_like("dgrad", dy, Wp, geom))]
    if os.environ.get("WAVE_CASES") == "step":
        bp = make_param(torch.randn(Cout, generator=g))
        xm = dev(torch.randn(N, H, H, Cin, generator=g))
        import numpy as np
        xm.bits = torch.randint(-32768, 32767, (N * H * H, (Cin + 15) // 16), dtype=torch.int16, device="cuda")
        variants += [("fprop+bias+lrelu+bits", lambda: E.conv_like("fprop", x, Wp, geom, bias=bp, act=K.ACT_LRELU, leak=0.2)),
                     ("dgrad+maskbits", lambda: E.conv_like("dgrad", dy, Wp, geom, out_mask=(xm, K.ACT_LRELU, 0.2)))]
    for name, fn in variants:
        if name == "dgrad" and k < 2:
            continue
        for _ in range(3): fn()
        t, tc = timed(fn)
        print("N%4d H%2d %4d->%4d k%d %-22s warm %.3f ms (%5.0f TF/s)   cold %.3f ms (%5.0f TF/s)" % (N, H, Cin, Cout, k, name, t, fl / t / 1e9, tc, fl / tc / 1e9), flush=True)
