"""GPU parity: each CUDA kernel family through the C ABI vs the CPU oracle (bit-for-bit inputs:
operands are bf16-rounded on both sides; accumulation is fp32 on the GPU, fp32 on the CPU)."""
import math

import pytest
import torch

from tests import parity as P
from b200gan import _capi as K
from b200gan import engine as E

pytestmark = pytest.mark.gpu

# bf16 output rounding is 2^-9 relative per element; fp32 outputs (wgrad) are much tighter
TOL = {"fprop": 6e-3, "dgrad": 6e-3, "wgrad": 2e-3}

CONV_CASES = [
    # N, H, W, Cin, Cout, k, stride          what it exercises
    (4, 16, 16, 64, 64, 5, 2),               # aligned tensor-core case
    (4, 16, 16, 200, 400, 5, 2),             # IWGAN c2 shape (ragged K chunk, N tile 208)
    (16, 8, 8, 400, 800, 5, 2),              # IWGAN c3 shape
    (3, 8, 8, 72, 40, 5, 2),                 # ragged everything, partial M tile
    (2, 7, 7, 64, 32, 5, 2),                 # odd spatial size: pad (2,2), deconv 4 -> 7
    (6, 4, 4, 256, 96, 1, 1),                # 1x1 conv (autoencoder c5)
    (64, 1, 1, 200, 512, 1, 1),              # dense as a 1x1 conv on 1x1 images
    (2, 32, 32, 64, 128, 4, 2),              # pix2pix k4 s2, pad (1,1)
    (2, 64, 64, 24, 16, 5, 2),               # wide rows: tile = part of one image
    (8, 32, 32, 3, 200, 5, 2),               # small-channel path (IWGAN c1 / dc-last)
    (4, 28, 28, 1, 64, 5, 2),                # MNIST-shaped first conv
    (2, 16, 16, 4, 64, 4, 2),                # pix2pix PatchGAN first conv (rgb+depth)
    (4, 64, 64, 4, 64, 4, 2),                # same, several tiles per image: fused fprop in its virtual-fifth-row form (k*Cin = 16)
    (16, 2, 2, 512, 512, 4, 2),              # pix2pix e8 / d1: 16 output pixels, K = 16 x 512 -> split-K over taps
    (16, 4, 4, 512, 512, 4, 2),              # pix2pix e7 / d2
    (16, 16, 16, 512, 512, 4, 2),            # pix2pix e5: 1024 output pixels, split-K
    (16, 4, 4, 1024, 512, 4, 2),             # d2's conv geometry with the concatenated 1024 channels (dgrad splits)
    (3, 8, 8, 136, 72, 5, 2),                # split-K with ragged K chunk, ragged N and a partial tile
    (128, 16, 16, 208, 400, 5, 2),           # stream-K schedule: 32 items over 74 CTA pairs, ~3 pieces per item
    (96, 8, 8, 400, 800, 5, 2),              # stream-K with a 32-wide K tail (not merged) and 4 N tiles
    (40, 32, 32, 64, 128, 4, 2),             # stream-K on a pix2pix mid layer (k4 s2, short items)
    (16, 128, 128, 64, 128, 4, 2),           # pix2pix e2 / d7: persistent 2-CTA kernel, whole items, double-buffered TMEM
    (8, 64, 64, 128, 256, 4, 2),             # pix2pix e3 / d6 (dgrad: N tile 128; fprop: N tile 256 -> plain kernel)
    (256, 16, 16, 64, 128, 5, 2),            # VAE c2 / dc3 at its BASELINE batch: persistent, ragged phases (9/6/6/4 taps)
    (512, 16, 16, 208, 400, 5, 2),           # bench.py's c2 exactly (B=512, padded 200 -> 208): 2-CTA schedules, stream-K
    (512, 8, 8, 400, 800, 5, 2),             # bench.py's c3 exactly
    (512, 32, 32, 3, 208, 5, 2),             # bench.py's c1 / last deconv exactly
    (16, 16, 16, 512, 1, 4, 2),              # pix2pix PatchGAN head m5 (one output channel): smallout1_* backward kernels
    (3, 9, 7, 64, 1, 5, 2),                  # the same kernels on odd sizes / k5 (borders, partial pixel blocks)
    (2, 8, 8, 24, 2, 4, 2),                  # two output channels: generic small-output backward kernels
]


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_family(case):
    res = P.conv_case(*case)
    for op, err in res.items():
        assert err < TOL[op], (case, res)


@pytest.mark.parametrize("case", [c for c in CONV_CASES if c[3] > 4 and c[0] <= 256],
                         ids=lambda c: "x".join(map(str, c)))
def test_conv_family_two_cta_layout_forced(case):
    """The planner gives layers whose 2-CTA work items cover < 65 % of the SMs single-tile CTAs (tapgemm_dual);
    with the threshold at 0 the same small cases run the two-tile / 2-CTA / persistent kernels the big layers use."""
    K.set_tuning("dual_min_pct", 0)
    try:
        res = P.conv_case(*case)
    finally:
        K.set_tuning("dual_min_pct", 65)
    for op, err in res.items():
        assert err < TOL[op], (case, res)


@pytest.mark.parametrize("case", [c for c in CONV_CASES if c[3] <= 4], ids=lambda c: "x".join(map(str, c)))
def test_small_channel_simt_fallback(case):
    """The image-side layers without a workspace: coalesced SIMT kernels instead of im2col + GEMM."""
    E.SMALL_CHANNEL_GEMM = False
    try:
        res = P.conv_case(*case)
    finally:
        E.SMALL_CHANNEL_GEMM = True
    for op, err in res.items():
        assert err < TOL[op], (case, res)


IMG_CASES = [
    # N, H, W, Cin, Cout, k, stride            fused image-side route (img_conv.cu): gather in the producer warps
    (8, 32, 32, 3, 208, 5, 2),                 # IWGAN c1 at the padded channel count, whole tiles
    (3, 20, 20, 3, 64, 5, 2),                  # M = 300: partial last tile, sign words stored from registers
    (2, 64, 64, 3, 64, 4, 2),                  # pix2pix first conv (k4: one 128-byte K chunk, no tail)
    (5, 12, 12, 3, 32, 3, 1),                  # k3 s1
    (4, 28, 28, 1, 64, 5, 2),                  # MNIST first conv (Cin 1: odd / even window starts)
    (2, 18, 18, 2, 48, 5, 2),                  # Cin 2
    (300, 32, 32, 3, 208, 5, 2),               # more tiles than SMs: the A ring and both accumulators wrap
]


@pytest.mark.parametrize("case", IMG_CASES, ids=lambda c: "x".join(map(str, c)))
def test_image_side_fused(case):
    """Fused gather fprop (bias in the spare K column, packed epilogue, sign words) and the filter gradient on the
    gathered rows with the bias gradient folded in, vs the oracle; then the same with a sign-bitmap mask."""
    from oracle import tf_ops as OT
    N, H, W, Cin, Cout, k, s = case
    E.begin()
    g = torch.Generator().manual_seed(11)
    geom = E.conv_geom(N, H, W, Cin, Cout, k, s)
    assert K.wgrad_folds_bias(geom, True), "case is meant for the fused route"
    x = P.bf16_round(torch.randn(N, H, W, Cin, generator=g))
    Wt = P.bf16_round(torch.randn(k, k, Cin, Cout, generator=g) / (k * (Cin ** 0.5)))
    b = torch.randn(Cout, generator=g) * 0.3
    Wp, bp = P.make_param(Wt, "w"), P.make_param(b, "b")
    pre_ref = OT.conv2d(x, Wt, b, s, None, None)
    y_ref = torch.where(pre_ref > 0, pre_ref, 0.2 * pre_ref)
    go = P.bf16_round(torch.randn(N, geom.Ho, geom.Wo, Cout, generator=g))
    with E.recording(True, active=[Wp, bp]):
        xt = P.dev(x)
        y = E.conv_like("fprop", xt, Wp, geom, bias=bp, act=K.ACT_LRELU, leak=0.2)
        torch.cuda.synchronize()
        yv = y.torch().float().cpu()
        assert P.rel_err(yv, y_ref) < 6e-3
        if y.bits is not None:
            # bit j of word (row, c/16) = out[row, c + j] > 0
            words = y.bits.cpu().to(torch.int32) & 0xffff
            got = torch.stack([(words >> j) & 1 for j in range(16)], dim=-1).reshape(words.shape[0], -1)[:, :Cout]
            assert torch.equal(got.bool(), (yv.reshape(-1, Cout) > 0))
        E.backward([(y, P.dev(go))], wrt=[])
    torch.cuda.synchronize()
    xr, Wr, br = x.clone(), Wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    gw_ref, gb_ref = torch.autograd.grad(OT.conv2d(xr, Wr, br, s, None, None), [Wr, br], go)
    assert P.rel_err(Wp.g32.reshape(Wt.shape), gw_ref) < 2e-3
    assert P.rel_err(bp.g32, gb_ref) < 2e-3
    # no bias, no activation, output multiplied by lrelu'(a) read from a's sign bitmap
    a_prev = E.conv_like("fprop", P.dev(x), Wp, geom, bias=bp, act=K.ACT_LRELU, leak=0.2)
    if a_prev.bits is not None:
        x2 = P.bf16_round(torch.randn(N, H, W, Cin, generator=g))
        z = E.conv_like("fprop", P.dev(x2), Wp, geom, out_mask=a_prev.mask)
        torch.cuda.synchronize()
        z_ref = OT.conv2d_same(x2, Wt, s) * P.act_grad_from_out(a_prev.torch().float().cpu(), K.ACT_LRELU)
        assert P.rel_err(z.torch().float(), z_ref) < 6e-3


IMG_DGRAD_CASES = [
    # N, H, W, Cin, Cout, k, stride            fused image-side dgrad: GEMM per image + col2im gather from shared memory
    (8, 32, 32, 3, 208, 5, 2),                 # IWGAN c1 input gradient / last deconv forward: 2 tiles per image, K tail 16
    (150, 32, 32, 3, 208, 5, 2),               # more images than SMs: TMEM buffers and the ring wrap
    (4, 16, 32, 3, 64, 5, 2),                  # 8x16 output pixels: 1 tile per image, no K tail
    (6, 16, 16, 3, 80, 4, 1),                  # stride 1, even filter
    (3, 32, 32, 1, 128, 5, 2),                 # Cin 1
    (5, 16, 16, 2, 144, 3, 1),                 # k3 s1, Cin 2, K = 2 chunks + tail
]


@pytest.mark.parametrize("case", IMG_DGRAD_CASES, ids=lambda c: "x".join(map(str, c)))
def test_image_side_fused_dgrad(case):
    """conv^T as GEMM + in-kernel col2im vs autograd of the oracle's conv: plain bf16, fp32 output (the gradient
    penalty's dx), bias + tanh (generator output layer), and a fused tanh' value mask (critic -> generator path)."""
    from oracle import tf_ops as OT
    N, H, W, Cin, Cout, k, s = case
    E.begin()
    g = torch.Generator().manual_seed(13)
    geom = E.conv_geom(N, H, W, Cin, Cout, k, s)
    Wt = P.bf16_round(torch.randn(k, k, Cin, Cout, generator=g) / (k * (Cout ** 0.5)))
    dy = P.bf16_round(torch.randn(N, geom.Ho, geom.Wo, Cout, generator=g))
    b = torch.randn(Cin, generator=g) * 0.3
    Wp, bp = P.make_param(Wt, "w"), P.make_param(b, "b")
    xr = torch.zeros(N, H, W, Cin, requires_grad=True)
    (gx_ref,) = torch.autograd.grad(OT.conv2d_same(xr, Wt, s), [xr], dy)
    gx = E.conv_like("dgrad", P.dev(dy), Wp, geom)
    gx32 = E.conv_like("dgrad", P.dev(dy), Wp, geom, out_f32=True)
    gt = E.conv_like("dgrad", P.dev(dy), Wp, geom, bias=bp, act=K.ACT_TANH)
    a_prev = P.bf16_round(torch.tanh(torch.randn(N, H, W, Cin, generator=g)))
    gm = E.conv_like("dgrad", P.dev(dy), Wp, geom, out_mask=(P.dev(a_prev), K.ACT_TANH, 0.0))
    torch.cuda.synchronize()
    assert P.rel_err(gx.torch().float(), gx_ref) < 6e-3
    assert P.rel_err(gx32.torch().float(), gx_ref) < 2e-3
    assert P.rel_err(gt.torch().float(), torch.tanh(gx_ref + b)) < 6e-3
    assert P.rel_err(gm.torch().float(), gx_ref * (1 - a_prev * a_prev)) < 6e-3


def test_conv_dgrad_fused_mask():
    res = P.conv_case(4, 16, 16, 200, 400, 5, 2, with_mask=True)
    assert res["dgrad"] < TOL["dgrad"], res
    res = P.conv_case(16, 4, 4, 512, 512, 4, 2, with_mask=True)          # split-K: the mask is applied by the finalize pass
    assert res["dgrad"] < TOL["dgrad"], res
    res = P.conv_case(128, 16, 16, 208, 400, 5, 2, with_mask=True)       # stream-K finisher applies the fused mask
    assert res["dgrad"] < TOL["dgrad"], res
    res = P.conv_case(4, 32, 32, 3, 200, 5, 2, with_mask=True)
    assert res["dgrad"] < TOL["dgrad"], res
    res = P.conv_case(16, 16, 16, 512, 1, 4, 2, with_mask=True)          # one-output-channel dgrad with the value mask
    assert res["dgrad"] < TOL["dgrad"], res


def test_gemv_outer_colsum():
    E.begin()
    g = torch.Generator().manual_seed(3)
    M, Kd = 48, 12800
    a = P.bf16_round(torch.randn(M, Kd, generator=g))
    w = P.bf16_round(torch.randn(Kd, 1, generator=g) / math.sqrt(Kd))
    b = torch.randn(1, generator=g)
    Wp, bp = P.make_param(w), P.make_param(b)
    out = E.dense_n1(P.dev(a), Wp, bp)
    torch.cuda.synchronize()
    assert P.rel_err(out.torch(), (a @ w + b).reshape(-1)) < 1e-5
    gvec = torch.randn(M, generator=g)
    gt = E.Tensor(gvec.cuda())
    like = P.dev(a); like.mask = (like, K.ACT_LRELU, 0.2)
    om = E.outer_mask(gt, Wp, like)
    want = gvec[:, None] * w.reshape(1, -1) * P.act_grad_from_out(a, K.ACT_LRELU)
    torch.cuda.synchronize()
    assert P.rel_err(om.torch().float(), want) < 6e-3
    E.launch("b200_colsum", E._p(like.buf), E._p(gt.buf), E._p(Wp.g32), M, Kd, 1.0)
    torch.cuda.synchronize()
    assert P.rel_err(Wp.g32, (gvec[:, None] * a).sum(0)) < 1e-5


@pytest.mark.parametrize("R,C", [(512, 12800), (8 * 8 * 8, 400), (37, 72), (1024, 3)])
def test_batch_norm_fwd_bwd(R, C):
    from oracle import tf_ops as OT
    E.begin()
    g = torch.Generator().manual_seed(5)
    z = P.bf16_round(torch.randn(R, C, generator=g) * 1.5 + 0.3)
    beta = torch.randn(C, generator=g) * 0.1
    bp = P.make_param(beta)
    with E.recording(True, active=[bp]):
        zt = P.dev(z); zt.requires_grad = True
        out = E.batch_norm_act(zt, bp, K.ACT_RELU, 0.0)
        zr = z.clone().requires_grad_(True); br = beta.clone().requires_grad_(True)
        ref = torch.relu(OT.batch_norm_train(zr, br))
        torch.cuda.synchronize()
        assert P.rel_err(out.torch().float(), ref) < 6e-3
        go = P.bf16_round(torch.randn(R, C, generator=g))
        gz_ref, gb_ref = torch.autograd.grad(ref, [zr, br], go)
        # deliverer applies relu' (mask convention)
        gdel = P.bf16_round(go * (ref > 0).float())
        (gz,) = E.backward([(out, P.dev(gdel))], wrt=[zt])
    torch.cuda.synchronize()
    assert P.rel_err(gz.torch().float(), gz_ref) < 1.5e-2
    assert P.rel_err(bp.g32, gb_ref) < 1e-3


def test_optimizers_match_tf_formulas():
    """Every branch of init_optimizer (util.py:150-183) through the fused update kernel vs the restated TF formulas;
    also the bf16 compute copy, the step counter and the in-pass gradient reset."""
    from oracle import tf_ops as OT
    E.begin()
    g = torch.Generator().manual_seed(7)
    n = 1000 + 64 * 3            # vector body + nothing left over (buckets are 64-aligned) ...
    cases = [(K.OPT_ADAM, "adam", 0.0), (K.OPT_RMSPROP, "rmsprop", 1.0), (K.OPT_SGD, "sgd", 0.0),
             (K.OPT_MOMENTUM, "momentum", 0.0), (K.OPT_ADAGRAD, "adagrad", 0.1), (K.OPT_ADADELTA, "adadelta", 0.0),
             (K.OPT_FTRL, "ftrl", 0.1), (K.OPT_CENTERED_RMSPROP, "centered", 1.0)]
    for n in (1192, 1001):       # ... and a ragged length for the scalar tail
        for kind, name, v0 in cases:
            p = torch.randn(n, generator=g); m = torch.zeros(n); v = torch.full((n,), v0); s3 = torch.zeros(n)
            dp, dm, dv, ds = p.cuda(), m.cuda(), v.cuda(), s3.cuda()
            p16 = torch.zeros(n, dtype=torch.bfloat16, device="cuda")
            step = torch.zeros(1, dtype=torch.int32, device="cuda")
            for t in range(1, 4):
                gr = torch.randn(n, generator=g)
                if name == "adam":
                    OT.adam_step(p, gr, m, v, t, 1e-3, 0.5, 0.9); args = (1e-3, 0.5, 0.9, 1e-8)
                elif name == "rmsprop":
                    OT.rmsprop_step(p, gr, v, m, 1e-3, 0.9, 0.01); args = (1e-3, 0.9, 0.01, 1e-10)
                elif name == "centered":
                    OT.centered_rmsprop_step(p, gr, v, s3, m, 1e-3, 0.9, 0.01); args = (1e-3, 0.9, 0.01, 1e-10)
                elif name == "sgd":
                    OT.sgd_step(p, gr, 1e-2); args = (1e-2, 0.0, 0.0, 0.0)
                elif name == "momentum":
                    OT.momentum_step(p, gr, m, 1e-2, 0.9); args = (1e-2, 0.9, 0.0, 0.0)
                elif name == "adagrad":
                    OT.adagrad_step(p, gr, v, 1e-2); args = (1e-2, 0.0, 0.0, 0.0)
                elif name == "adadelta":
                    OT.adadelta_step(p, gr, v, m, 1.0); args = (1.0, 0.95, 0.0, 1e-8)
                else:
                    OT.ftrl_step(p, gr, v, m, 1e-2); args = (1e-2, 0.0, 0.0, 0.0)
                dg = gr.cuda()
                E.launch("b200_optim_step", E._p(dp), E._p(dm), E._p(dv), E._p(ds), E._p(dg), E._p(p16), n, kind, *args,
                         1.0, 0.0, 1, E._p(step))
                torch.cuda.synchronize()
                assert float(dg.abs().max()) == 0.0, name                 # gradient reset in the same pass
            assert torch.allclose(dp.cpu(), p, rtol=3e-5, atol=2e-6), (name, n)
            assert int(step.item()) == 3
            assert torch.equal(p16.cpu(), dp.cpu().to(torch.bfloat16))


def test_transpose_batch_matches_per_tensor_transposes():
    """The one-launch re-layout of a group's K-major weight copies (b200_transpose_batch) == per-tap transposes."""
    import argparse
    from b200gan import session as S
    from b200gan.models import gan as gan_model
    sess = S.Session(seed=0)
    x_in = S.Input(4, (32, 32, 3), slots=2)
    args = argparse.Namespace(model="iwgan", batch_size=4, latent_size=24, n_disc_train=1, optimizer="adam", lr=1e-4,
                              beta1=0.5, beta2=0.9)
    gan_model.gan(x_in, args)
    sess.begin_step()
    n = 0
    for grp in sess.store.groups:
        grp.p16.copy_(torch.randn(grp.size, device="cuda"))
        grp.refresh_transposed()
        torch.cuda.synchronize()
        for p in grp.params:
            if p.p16_t is None:
                continue
            a, b = p.shape[-2], p.shape[-1]
            want = p.p16.reshape(-1, a, b).transpose(1, 2).contiguous().reshape(-1)
            assert torch.equal(p.p16_t, want), p.name
            n += 1
    assert n >= 3


def test_philox_moments_and_counter():
    E.begin()
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    a = E.random_fill((1 << 20,), True, 1234, cnt, 0)
    b = E.random_fill((1 << 20,), True, 1234, cnt, 0)
    u = E.random_fill((1 << 20,), False, 1234, cnt, 1)
    torch.cuda.synchronize()
    assert int(cnt.item()) == 3
    assert abs(float(a.torch().mean())) < 5e-3 and abs(float(a.torch().std()) - 1) < 5e-3
    assert not torch.equal(a.torch(), b.torch())
    assert 0.0 <= float(u.torch().min()) and float(u.torch().max()) < 1.0 and abs(float(u.torch().mean()) - 0.5) < 3e-3


def test_interp_sumsq_wgan_loss():
    E.begin()
    g = torch.Generator().manual_seed(9)
    B, D = 16, 3072
    x = P.bf16_round(torch.rand(B, D, generator=g) * 2 - 1); f = P.bf16_round(torch.rand(B, D, generator=g) * 2 - 1)
    al = torch.rand(B, 1, generator=g)
    it = E.interpolate(P.dev(x), P.dev(f), E.Tensor(al.cuda()))
    torch.cuda.synchronize()
    assert P.rel_err(it.torch().float(), x + al * (f - x)) < 4e-3
    gr = torch.randn(B, D, generator=g)
    ss = E.sumsq(E.Tensor(gr.cuda()))
    dr, df = torch.randn(B, generator=g), torch.randn(B, generator=g)
    gl, dl = E.wgan_losses(E.Tensor(dr.cuda()), E.Tensor(df.cuda()), ss, 10.0)
    torch.cuda.synchronize()
    s = float((gr.double() ** 2).sum().sqrt())
    assert abs(float(gl.buf.item()) + float(df.mean())) < 1e-5
    want = float(df.mean() - dr.mean()) + 10.0 * (s - 1) ** 2
    assert abs(float(dl.buf.item()) - want) < 1e-3 * abs(want)


def test_sign_bitmaps_match_value_masks():
    """relu/lrelu epilogues also write a 1-bit/element sign map; a gradient epilogue masked through the
    bitmap must equal the one masked through the stored activation, bit for bit."""
    import numpy as np
    E.begin()
    g = torch.Generator().manual_seed(3)
    N = 6
    chain = [(32, 3, 200), (16, 200, 400), (8, 400, 800)]          # IWGAN critic c1, c2, c3
    x = P.dev(torch.randn(N, 32, 32, 3, generator=g))
    acts = []
    E.TAP_SPLIT = False            # (at this tiny batch the launches would otherwise take the split-K route: no bitmaps)
    for (H, Cin, Cout) in chain:
        geom = E.conv_geom(N, H, H, Cin, Cout, 5, 2)
        Wp = P.make_param(torch.randn(5, 5, Cin, Cout, generator=g) * (0.3 / math.sqrt(25 * Cin)) * 5)
        bp = P.make_param(torch.randn(Cout, generator=g) * 0.1)
        y = E.conv_like("fprop", x, Wp, geom, bias=bp, act=K.ACT_LRELU, leak=0.2)
        assert y.bits is not None, (H, Cin, Cout)
        torch.cuda.synchronize()
        words = y.bits.cpu().numpy().view(np.uint16)                # [rows, ceil(C/16)]
        unpacked = ((words[:, :, None] >> np.arange(16, dtype=np.uint16)) & 1).reshape(words.shape[0], -1)[:, :Cout]
        ref = (y.torch().float().cpu().numpy().reshape(-1, Cout) > 0)
        assert (unpacked.astype(bool) == ref).all(), (H, Cin, Cout)
        acts.append((x, y, geom, Wp))
        x = y
    for (xin, y, geom, Wp) in acts[1:]:
        dy = P.dev(torch.randn(*y.shape, generator=g))
        via_bits = E.conv_like("dgrad", dy, Wp, geom, out_mask=(xin, K.ACT_LRELU, 0.2))
        saved, xin.bits = xin.bits, None
        via_vals = E.conv_like("dgrad", dy, Wp, geom, out_mask=(xin, K.ACT_LRELU, 0.2))
        xin.bits = saved
        torch.cuda.synchronize()
        assert torch.equal(via_bits.torch(), via_vals.torch())
        # double-backward shape: fprop masked by the consumer's activation
        v = P.dev(torch.randn(*xin.shape, generator=g))
        f_bits = E.conv_like("fprop", v, Wp, geom, out_mask=(y, K.ACT_LRELU, 0.2))
        saved, y.bits = y.bits, None
        f_vals = E.conv_like("fprop", v, Wp, geom, out_mask=(y, K.ACT_LRELU, 0.2))
        y.bits = saved
        torch.cuda.synchronize()
        assert torch.equal(f_bits.torch(), f_vals.torch())
    E.TAP_SPLIT = True
