#!/usr/bin/env python
"""Achieved HBM bandwidth of the memory-bound kernels of the IWGAN-32 step (B=512, L=200) through the C ABI.

For every kernel: ALGORITHMIC bytes (what the op must read + write once) / CUDA-event time, L2
flushed before every timed launch (a 256 MB write), median of `--reps` launches, against MEASURED_PEAKS.json's
copy bandwidth.  Writes profiles-style CSV to --out.

  python tools/bench_membound.py --out gpurun_out/membound.csv
"""
import argparse
import csv
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/membound.csv")
    ap.add_argument("--reps", type=int, default=7)
    a = ap.parse_args()
    import torch
    import b200gan  # noqa: F401
    from b200gan import _capi as K
    from b200gan import engine as E
    E.begin()
    peak = 6547.2
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk)).get("hbm_gbs", peak))
    dev = "cuda"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    bf, f32 = torch.bfloat16, torch.float32
    P = E._p

    def rnd(shape, dtype=bf):
        return torch.randn(shape, device=dev, dtype=torch.float32).to(dtype)

    rows = []

    def timed(name, shape, nbytes, fn, note=""):
        ts = []
        for _ in range(a.reps + 2):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        us = statistics.median(ts[2:])
        gbs = nbytes / us * 1e-3
        rows.append([name, shape, nbytes, round(us, 2), round(gbs, 1), round(gbs / peak, 3), note])
        print("%-28s %-26s %10.1f MB %9.1f us %8.1f GB/s  %.2f of peak  %s" % (name, shape, nbytes / 1e6, us, gbs, gbs / peak, note))

    L = lambda name, *args, **kw: E.launch(name, *args, **kw)

    # ---- optimizer: critic group (10.03 M parameters) and generator group (12.58 M)
    for n in (10_032_448, 12_582_912):
        p, m, v, g = (torch.randn(n, device=dev) for _ in range(4))
        v.abs_()
        p16 = torch.empty(n, device=dev, dtype=bf)
        step = torch.zeros(1, dtype=torch.int32, device=dev)
        timed("optim_kernel(adam)", "n=%d" % n, n * 34,
              lambda: L("b200_optim_step", P(p), P(m), P(v), None, P(g), P(p16), n, K.OPT_ADAM, 1e-4, 0.5, 0.9, 1e-8,
                        1.0, 0.0, 1, P(step)),
              "read g,p,m,v; write p,m,v,g=0,bf16 copy (34 B/param)")
    # ---- batch norm of the generator layers: fc1 [512,12800], dc1 [512*8*8,400], dc2 [512*16*16,208]
    for R, C in ((512, 12800), (512 * 64, 400), (512 * 256, 208)):
        z = rnd((R, C)); out = torch.empty_like(z); gz = rnd((R, C)); dz = torch.empty_like(z)
        stats = torch.zeros(2 * C, device=dev); beta = torch.zeros(C, device=dev); bsum = torch.zeros(2 * C, device=dev)
        timed("bn_sums(colsum_vec<1>)", "%dx%d" % (R, C), R * C * 2, lambda: L("b200_bn_sums", P(z), P(stats), R, C), "1 read")
        timed("bn_apply_vec", "%dx%d" % (R, C), R * C * 4,
              lambda: L("b200_bn_apply", P(z), P(stats), P(beta), P(out), R, C, 1e-3, K.ACT_RELU, 0.0), "1 read + 1 write")
        timed("bn_bwd(sums+apply)", "%dx%d" % (R, C), R * C * 10,
              lambda: L("b200_bn_bwd", P(gz), P(z), P(stats), P(bsum), P(dz), R, C, 1e-3), "2x(g,z) reads + 1 write")
    # ---- bias-gradient column sums of the critic convs (dY of c1, c2, c3) and the fc2 weight gradient
    for R, C in ((512 * 256, 208), (512 * 64, 400), (512 * 16, 800)):
        x = rnd((R, C)); o = torch.zeros(C, device=dev)
        timed("colsum_vec<0>", "%dx%d" % (R, C), R * C * 2, lambda: L("b200_colsum", P(x), None, P(o), R, C, 1.0), "1 read")
    x = rnd((512, 12800)); wrow = torch.randn(512, device=dev); o = torch.zeros(12800, device=dev)
    timed("colsum_vec<0>(weighted)", "512x12800", 512 * 12800 * 2, lambda: L("b200_colsum", P(x), P(wrow), P(o), 512, 12800, 1.0), "1 read")
    # ---- critic head
    w16 = rnd((12800,)); b1 = torch.zeros(1, device=dev); o = torch.empty(512, device=dev); o2 = torch.empty_like(x)
    timed("gemv_rows", "512x12800", 512 * 12800 * 2, lambda: L("b200_gemv_rows", P(x), P(w16), P(b1), P(o), 512, 12800, 0, 0.0), "1 read")
    timed("outer_mask_vec", "512x12800", 512 * 12800 * 4,
          lambda: L("b200_outer_mask", P(wrow), P(w16), P(x), P(o2), 512, 12800, K.ACT_LRELU, 0.2), "1 read (mask) + 1 write")
    # ---- gradient penalty helpers on the [512, 3072] images
    xi = rnd((512, 3072)); gi = rnd((512, 3072)); al = torch.rand(512, device=dev); oi = torch.empty_like(xi)
    gf = torch.randn(512, 3072, device=dev); ss = torch.zeros(1, device=dev)
    timed("interp", "512x3072", 512 * 3072 * 6, lambda: L("b200_interp", P(xi), P(gi), P(al), P(oi), 512, 3072), "2 reads + 1 write (latency-bound: 9 MB)")
    timed("reduce_sum(sumsq fp32)", "512x3072", 512 * 3072 * 4, lambda: L("b200_reduce_sum", P(gf), 1, 512 * 3072, P(ss), 1.0, 1), "1 read (6 MB)")
    u8 = torch.randint(0, 256, (512, 32, 32, 3), device=dev, dtype=torch.uint8)
    timed("affine_act(u8 -> bf16)", "512x32x32x3", 512 * 3072 * 3, lambda: L("b200_affine_act", P(u8), 2, P(oi), 0, 512 * 3072, 2 / 255.0, -1.0, 0, 0.0), "input stage: 1 B read + 2 B write")
    # ---- weight re-layout: c2 + c3 tap transposes of the critic group
    src = rnd((25 * 208 * 400 + 25 * 400 * 800,)); dst = torch.empty_like(src)
    ents = (K.TransposeEntry * 2)()
    tiles = 0
    off = 0
    for e, (t, aa, bb) in zip(ents, ((25, 208, 400), (25, 400, 800))):
        e.src, e.dst, e.tile_begin, e.T, e.A, e.B = src.data_ptr() + off * 2, dst.data_ptr() + off * 2, tiles, t, aa, bb
        tiles += t * ((aa + 63) // 64) * ((bb + 63) // 64)
        off += t * aa * bb
    tab = torch.frombuffer(bytearray(bytes(ents)), dtype=torch.uint8).clone().to(dev)
    timed("transpose_batch", "c2+c3 taps (10.08 M)", off * 4, lambda: L("b200_transpose_batch", P(tab), 2, tiles), "1 read + 1 write (bf16)")
    # ---- concat slice copy (pix2pix d2: 512 + 512 channels at 2x2... use the 128x128x(64+64) skip, B=16)
    a1 = rnd((16 * 128 * 128, 64)); cat = torch.empty((16 * 128 * 128, 128), device=dev, dtype=bf)
    timed("slice_cols<8>", "262144x64 -> x128", 262144 * 64 * 4, lambda: L("b200_slice_cols", P(a1), 64, 0, P(cat), 128, 64, 262144, 64, None, 0, 0.0), "1 read + 1 write")
    # ---- elementwise gradient mask
    gm = rnd((512 * 64, 400)); am = rnd((512 * 64, 400)); om = torch.empty_like(gm)
    timed("maskmul", "32768x400", 32768 * 400 * 6, lambda: L("b200_maskmul", P(gm), P(am), P(om), 32768 * 400, K.ACT_LRELU, 0.2), "2 reads + 1 write")

    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    with open(a.out, "w", newline="") as f:
        wri = csv.writer(f)
        wri.writerow(["# achieved HBM GB/s = algorithmic bytes / CUDA-event time, L2 flushed before each launch, median of %d; peak %.1f GB/s (MEASURED_PEAKS.json)" % (a.reps, peak)])
        wri.writerow(["kernel", "shape", "algorithmic_bytes", "us", "GB/s", "frac_of_hbm_peak", "note"])
        wri.writerows(rows)


if __name__ == "__main__":
    main()
