"""Planner sweep: every tensor-core conv geometry of the named workloads (fprop and dgrad) under the tap-GEMM
planner's degrees of freedom -- two-tile/2-CTA layout or single-tile CTAs, N-tile cap, split-K factor -- forced
through b200_set_tuning.  Prints, per geometry, the planner's own choice ("auto") beside every forced combination;
the heuristics in csrc/capi.cu (pick_bn_tile_conv, tc_tap_splits) and tc_gemm.cu (tapgemm_dual) are fitted to it.
Timing: 10 launches captured in a CUDA graph, median of 5 replays (L2 warm: these layers' operands fit the L2)."""
import os, sys, json
import ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200gan  # noqa
from b200gan import engine as E, _capi as K
from tests.parity import make_param, dev

E.begin()
SETS = {
    "p2p": [(16, 128, 64, 128, 4), (16, 64, 128, 256, 4), (16, 32, 256, 512, 4), (16, 16, 512, 512, 4),
            (16, 8, 512, 512, 4), (16, 4, 512, 512, 4), (16, 2, 512, 512, 4), (16, 128, 64, 256, 4), (16, 64, 128, 512, 4),
            (16, 32, 256, 1024, 4), (16, 16, 512, 1024, 4), (16, 8, 512, 1024, 4), (16, 4, 512, 1024, 4)],
    "vae": [(256, 16, 64, 128, 5), (256, 8, 128, 256, 5), (256, 4, 256, 256, 5)],
    "cnn": [(64, 14, 64, 128, 5), (64, 7, 128, 256, 5), (64, 4, 256, 256, 5)],
    "iwgan": [(512, 16, 208, 416, 5), (512, 8, 416, 832, 5)],
}
which = [w for w in os.environ.get("TUNE_SETS", "p2p,vae,cnn").split(",") if w]


def timed(fn, inner=10, reps=5):
    st = torch.cuda.Stream()
    E.S.stream = C.c_void_p(st.cuda_stream)
    keep = []
    with torch.cuda.stream(st):
        fn(); torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=st):
            for _ in range(inner):
                keep.append(fn())
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / inner * 1e3)
        ts.sort()
    return ts[len(ts) // 2]


def set_all(dual_pct, bn_cap, splits):
    K.set_tuning("dual_min_pct", dual_pct)
    K.set_tuning("bn_tile_cap", bn_cap)
    K.set_tuning("tap_splits", splits)


out = {}
for name in which:
    for (N, H, Cin, Cout, k) in SETS[name]:
        g0 = torch.Generator().manual_seed(0)
        geom = E.conv_geom(N, H, H, Cin, Cout, k, 2)
        x = dev(torch.randn(N, H, H, Cin, generator=g0)); dy = dev(torch.randn(N, geom.Ho, geom.Wo, Cout, generator=g0))
        Wp = make_param(torch.randn(k, k, Cin, Cout, generator=g0) * 0.05)
        bp = make_param(torch.randn(Cout, generator=g0))
        fl = 2.0 * N * geom.Ho * geom.Wo * k * k * Cin * Cout
        for op, fn in (("fprop", lambda: E.conv_like("fprop", x, Wp, geom, bias=bp, act=K.ACT_LRELU, leak=0.2)),
                       ("dgrad", lambda: E.conv_like("dgrad", dy, Wp, geom))):
            set_all(65, 0, -1)
            auto = timed(fn)
            res = []
            for dual_pct in (0, 100000):
                for bn_cap in (256, 128, 64):
                    for splits in (1, 2, 4, 8):
                        set_all(dual_pct, bn_cap, splits)
                        try:
                            t = timed(fn)
                        except Exception as ex:        # a combination the kernels reject
                            t = float("inf")
                        res.append((t, "dual" if dual_pct == 0 else "single", bn_cap, splits))
            set_all(65, 0, -1)
            res.sort()
            key = "%s N%d %dx%dx%d->%d k%d %s" % (name, N, H, H, Cin, Cout, k, op)
            out[key] = {"auto_us": auto, "best": res[:4]}
            print("%-44s auto %6.1f us (%5.0f TF/s) | best %s" % (
                key, auto, fl / auto / 1e6, "  ".join("%.1f:%s/bn%d/s%d" % r for r in res[:4])), flush=True)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
