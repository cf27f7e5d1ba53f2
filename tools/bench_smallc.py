"""Micro-benchmark of the image-side (<= 4 channel) conv route: fprop / dgrad / wgrad of IWGAN c1 and the
generator's last deconv, 10 back-to-back launches in a CUDA graph."""
import os, sys
import ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200gan  # noqa
from b200gan import engine as E, _capi as K
from tests.parity import make_param, dev

E.begin()

def timed(fn, reps=5, inner=10):
    st = torch.cuda.Stream()
    E.S.stream = C.c_void_p(st.cuda_stream)
    keep = []
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        fn(); torch.cuda.synchronize()
        with torch.cuda.graph(graph, stream=st):
            for _ in range(inner):
                keep.append(fn())
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / inner)
    ts.sort()
    return ts[len(ts) // 2]

N, H, Cin, Cout, k = 512, 32, 3, int(os.environ.get("COUT", "208")), 5
g = torch.Generator().manual_seed(0)
geom = E.conv_geom(N, H, H, Cin, Cout, k, 2)
x = dev(torch.randn(N, H, H, Cin, generator=g)); dy = dev(torch.randn(N, geom.Ho, geom.Wo, Cout, generator=g))
Wp = make_param(torch.randn(k, k, Cin, Cout, generator=g) * 0.05); bp = make_param(torch.randn(Cout, generator=g))
def wgrad():
    ws, wsb = E._workspace(geom, 2)
    E.launch("b200_conv2d_wgrad", E._p(x.buf), E._p(dy.buf), E._p(Wp.g32), C.byref(geom), 1.0, E._p(ws), wsb, 0)
    return ws
cases = {
    "fprop bias+lrelu": lambda: E.conv_like("fprop", x, Wp, geom, bias=bp, act=K.ACT_LRELU, leak=0.2),
    "fprop plain": lambda: E.conv_like("fprop", x, Wp, geom),
    "dgrad bf16": lambda: E.conv_like("dgrad", dy, Wp, geom),
    "dgrad f32": lambda: E.conv_like("dgrad", dy, Wp, geom, out_f32=True),
    "wgrad": wgrad,
}
for name, fn in cases.items():
    t = timed(fn)
    print("c1 %-18s %.1f us" % (name, t * 1e3), flush=True)
