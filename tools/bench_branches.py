"""Micro-benchmark: do independent layer chains on parallel graph branches recover the idle SMs of the
128-CTA tap GEMMs?  Three chains (the critic's D(real) / D(fake) / D(x_hat)) of
c2 fprop -> c3 fprop -> c3 dgrad -> c2 dgrad (+ optionally the two wgrads), captured in one CUDA graph either
back to back on one stream or forked onto three streams and joined."""
import os, sys
import ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200gan  # noqa
from b200gan import engine as E, _capi as K
from tests.parity import make_param, dev

E.begin()
N = int(os.environ.get("BR_N", 512))
WG = int(os.environ.get("BR_WGRAD", 0))
g0 = torch.Generator().manual_seed(0)
geo2 = E.conv_geom(N, 16, 16, 208, 416, 5, 2)     # physical (padded) channel counts of the step
geo3 = E.conv_geom(N, 8, 8, 416, 832, 5, 2)
W2 = make_param(torch.randn(5, 5, 208, 416, generator=g0) * 0.05)
W3 = make_param(torch.randn(5, 5, 416, 832, generator=g0) * 0.05)
b2 = make_param(torch.randn(416, generator=g0))
b3 = make_param(torch.randn(832, generator=g0))
xs = [dev(torch.randn(N, 16, 16, 208, generator=g0)) for _ in range(3)]
dys = [dev(torch.randn(N, 4, 4, 832, generator=g0)) for _ in range(3)]


def chain(i, keep):
    a2 = E.conv_like("fprop", xs[i], W2, geo2, bias=b2, act=K.ACT_LRELU, leak=0.2)
    a3 = E.conv_like("fprop", a2, W3, geo3, bias=b3, act=K.ACT_LRELU, leak=0.2)
    d2 = E.conv_like("dgrad", dys[i], W3, geo3, out_mask=a2.mask)
    d1 = E.conv_like("dgrad", d2, W2, geo2)
    keep += [a2, a3, d2, d1]
    if WG:
        for (a_, b_, W_, g_) in ((a2, dys[i], W3, geo3), (xs[i], d2, W2, geo2)):
            ws, wsb = E._workspace(g_, 2)
            E.launch("b200_conv2d_wgrad", E._p(a_.buf), E._p(b_.buf), E._p(W_.g32), C.byref(g_), 1.0, E._p(ws), wsb, 0)


def run(parallel, inner=4):
    main = torch.cuda.Stream()
    side = [torch.cuda.Stream() for _ in range(2)]
    keep = []

    def body():
        for _ in range(inner):
            if not parallel:
                for i in range(3):
                    chain(i, keep)
            else:
                for s in side:
                    s.wait_stream(main)
                for i in range(3):
                    st = main if i == 0 else side[i - 1]
                    with torch.cuda.stream(st):
                        E.S.stream = C.c_void_p(st.cuda_stream)
                        chain(i, keep)
                E.S.stream = C.c_void_p(main.cuda_stream)
                for s in side:
                    main.wait_stream(s)

    with torch.cuda.stream(main):
        E.S.stream = C.c_void_p(main.cuda_stream)
        body(); torch.cuda.synchronize(); keep.clear()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=main):
            body()
        ts = []
        for _ in range(7):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / inner)
        ts.sort()
    return ts[len(ts) // 2]


for rep in range(2):
    a = run(False)
    b = run(True)
    print("N=%d wgrad=%d  3 chains serial %.3f ms   3 branches %.3f ms   ratio %.3f" % (N, WG, a, b, b / a), flush=True)
