"""GPU parity of the Gen-2 (hem) layer options and the boundary helpers: dropout, instance norm, VALID padding,
residual blocks, NCHW / uint8 input stage, summaries (hem/ops/layers.py, hem/ops/images.py, ops/summaries.py)."""
import argparse
import math

import numpy as np
import pytest
import torch

from tests import parity as P
from b200gan import _capi as K
from b200gan import engine as E
from b200gan import session as S
from oracle import tf_ops as OT

pytestmark = pytest.mark.gpu


def test_dropout_forward_and_gradient():
    E.begin()
    g = torch.Generator().manual_seed(0)
    x = P.bf16_round(torch.randn(64, 8, 8, 48, generator=g))
    u = torch.rand(64, 8, 8, 48, generator=g)
    go = P.bf16_round(torch.randn(64, 8, 8, 48, generator=g))
    with E.recording(True):
        xt = P.dev(x); xt.requires_grad = True
        y = E.dropout(xt, 0.7, E.Tensor(u.cuda()))
        (gx,) = E.backward([(y, P.dev(go))], wrt=[xt])
    torch.cuda.synchronize()
    xr = x.clone().requires_grad_(True)
    ref = OT.dropout(xr, 0.7, u)
    (gref,) = torch.autograd.grad(ref, xr, go)
    assert P.rel_err(y.torch().float(), ref) < 6e-3
    assert P.rel_err(gx.torch().float(), gref) < 6e-3
    kept = float((y.torch() != 0).float().mean())
    assert abs(kept - 0.7) < 0.02


@pytest.mark.parametrize("N,H,C", [(4, 16, 64), (3, 7, 40), (16, 2, 512)])
def test_instance_norm_forward_backward(N, H, C):
    E.begin()
    g = torch.Generator().manual_seed(1)
    x = P.bf16_round(torch.randn(N, H, H, C, generator=g) * 1.3 + 0.2)
    sc, sh = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    go = P.bf16_round(torch.randn(N, H, H, C, generator=g))
    scp, shp = P.make_param(sc, "scale"), P.make_param(sh, "shift")
    with E.recording(True, active=[scp, shp]):
        xt = P.dev(x); xt.requires_grad = True
        y = E.instance_norm(xt, scp, shp)
        (gx,) = E.backward([(y, P.dev(go))], wrt=[xt])
    torch.cuda.synchronize()
    xr, scr, shr = x.clone().requires_grad_(True), sc.clone().requires_grad_(True), sh.clone().requires_grad_(True)
    ref = OT.instance_norm(xr, scr, shr)
    gxr, gsc, gsh = torch.autograd.grad(ref, [xr, scr, shr], go)
    assert P.rel_err(y.torch().float(), ref) < 6e-3
    assert P.rel_err(gx.torch().float(), gxr) < 1.5e-2
    assert P.rel_err(scp.g32, gsc) < 5e-3 and P.rel_err(shp.g32, gsh) < 5e-3


@pytest.mark.parametrize("case", [(4, 17, 17, 64, 96, 5, 2), (2, 32, 32, 64, 128, 4, 2), (8, 9, 9, 128, 64, 3, 1)],
                         ids=lambda c: "x".join(map(str, c)))
def test_valid_padding_conv_family(case):
    """padding='VALID' (hem/ops/layers.py:118,189): no padding, Ho = (H - k) // s + 1; fprop, dgrad and wgrad."""
    N, H, W, Cin, Cout, k, s = case
    E.begin()
    g = torch.Generator().manual_seed(2)
    x = P.bf16_round(torch.randn(N, H, W, Cin, generator=g))
    Wt = P.bf16_round(torch.randn(k, k, Cin, Cout, generator=g) / (k * math.sqrt(Cin)))
    geom = E.conv_geom(N, H, W, Cin, Cout, k, s, 'VALID')
    assert (geom.Ho, geom.pad_t) == ((H - k) // s + 1, 0)
    dy = P.bf16_round(torch.randn(N, geom.Ho, geom.Wo, Cout, generator=g))
    Wp = P.make_param(Wt)
    y = E.conv_like("fprop", P.dev(x), Wp, geom)
    gx = E.conv_like("dgrad", P.dev(dy), Wp, geom)
    xd, dyd = P.dev(x), P.dev(dy)
    E.launch("b200_conv2d_wgrad", E._p(xd.buf), E._p(dyd.buf), E._p(Wp.g32), E.C.byref(geom), 1.0, None, 0, 0)
    torch.cuda.synchronize()
    xr, Wr = x.clone().requires_grad_(True), Wt.clone().requires_grad_(True)
    ref = OT.conv2d_valid(xr, Wr, s)
    gxr, gwr = torch.autograd.grad(ref, [xr, Wr], dy)
    assert P.rel_err(y.torch().float(), ref) < 6e-3
    assert P.rel_err(gx.torch().float(), gxr) < 6e-3
    assert P.rel_err(Wp.g32.reshape(Wt.shape), gwr) < 2e-3


def _hem_session(B, shape, dtype=torch.float32, layout="NHWC"):
    sess = S.Session(seed=0)
    sess.use_graphs = False
    return sess, S.Input(B, shape, slots=1, dtype=dtype, layout=layout)


def test_residual_block_with_instance_norm_and_dropout_matches_oracle():
    """hem.residual (hem/ops/layers.py:216-320) with use_instance_norm and dropout, tanh activation (smooth: tight bar):
    every variable's gradient vs the oracle's composition of the same ops."""
    from collections import OrderedDict
    from b200gan.hem_ops import layers as hem
    from b200gan.ops.activations import tanh
    from b200gan.ops.layers import variable_scope
    from b200gan.variables import optimizer_cfg
    B, H, C = 8, 16, 32
    sess, x_in = _hem_session(B, (H, H, 3))
    args = argparse.Namespace(optimizer="adam", lr=1e-3, beta1=0.9, beta2=0.999)

    def net(batch01):
        with E.recording(True, active='all'), variable_scope('net'):
            x = hem.rescale(batch01, (0, 1), (-1, 1))
            h = hem.conv2d(x, 3, C, 3, 1, activation=tanh, name='in')
            h = hem.residual(h, C, C, 3, 1, use_instance_norm=True, activation=tanh, dropout=0.8, name='r')
            y = hem.conv2d(h, C, 3, 3, 1, activation=tanh, name='out')
            return E.eltloss(y, x, 5, scale=1.0 / y.numel)

    with sess.building():
        sess.store.begin_pass()
        E.backward([(net(x_in.next()), None)])
    x_in.materialize(sess.device)
    sess.store.finalize([('all', list(sess.store.params.values()), optimizer_cfg(args))], sess.device)
    gen = torch.Generator().manual_seed(4)
    p = OrderedDict()
    for n_, prm in sess.store.params.items():
        p[n_] = (torch.ones(prm.logical_shape) if n_.endswith('/scale') else
                 torch.zeros(prm.logical_shape) if n_.endswith('/shift') else
                 P.bf16_round(OT.xavier_uniform(prm.logical_shape, gen)))
    sess.store.load(p)
    x01 = P.bf16_round(torch.rand(B, H, H, 3, generator=gen))
    u1, u2 = torch.rand(B, H, H, C, generator=gen), torch.rand(B, H, H, C, generator=gen)
    x_in.feed(0, x01.cuda())
    sess.begin_step(); x_in.reset(); sess.store.groups[0].zero_grad()
    sess.noise_queue = [u1.clone(), u2.clone()]
    loss = net(x_in.next())
    E.backward([(loss, None)])
    torch.cuda.synchronize()
    q = OrderedDict((k_, v.clone().requires_grad_(True)) for k_, v in p.items())
    x = 2 * x01 - 1
    h = torch.tanh(OT.conv2d_same(x, q['net/vars/in/weights'], 1) + q['net/vars/in/bias'])
    sc = OT.conv2d_same(h, q['net/vars/rA/weights'], 1) + q['net/vars/rA/bias']
    a = OT.dropout(torch.tanh(OT.instance_norm(sc, q['net/vars/r/scale'], q['net/vars/r/shift'])), 0.8, u1)
    b = OT.instance_norm(OT.conv2d_same(a, q['net/vars/rB/weights'], 1) + q['net/vars/rB/bias'],
                         q['net/vars/r/scale'], q['net/vars/r/shift'])
    h2 = OT.dropout(torch.tanh(b + sc), 0.8, u2)
    y = torch.tanh(OT.conv2d_same(h2, q['net/vars/out/weights'], 1) + q['net/vars/out/bias'])
    ref = torch.mean((y - x) ** 2)
    grads = torch.autograd.grad(ref, list(q.values()))
    assert abs(float(loss.buf.item()) - float(ref)) < 2e-3 * abs(float(ref)) + 1e-5
    assert set(q) == set(sess.store.params)
    for (name, prm), want in zip(sess.store.params.items(), grads):
        if name == 'net/vars/rB/bias':            # feeds an instance norm: analytically zero gradient, ours is rounding noise
            assert float(prm.logical(prm.g32).abs().max()) < 1e-3 and float(want.abs().max()) < 1e-5
            continue
        e = P.rel_err(prm.logical(prm.g32), want)
        print("  [residual] %-24s err %.3e" % (name, e))
        assert e < 2e-2, (name, e)


def test_nchw_uint8_input_equals_nhwc_float_input():
    """session.Input(layout='NCHW', dtype=uint8): the reference's Gen-2 batches (NCHW) as decoded bytes are transposed,
    normalised and rescaled by hem.rescale in one pass; same result as the NHWC float path."""
    from b200gan.hem_ops import layers as hem
    B, H, C = 4, 32, 3
    gen = torch.Generator().manual_seed(5)
    raw = torch.randint(0, 256, (B, C, H, H), generator=gen, dtype=torch.uint8)
    outs = []
    for dtype, layout in ((torch.uint8, "NCHW"), (torch.float32, "NHWC")):
        sess, x_in = _hem_session(B, (C, H, H) if layout == "NCHW" else (H, H, C), dtype, layout)
        x_in.materialize(sess.device)
        x_in.feed(0, raw.cuda() if layout == "NCHW" else (raw.float() / 255).permute(0, 2, 3, 1).contiguous().cuda())
        sess.begin_step()
        y = hem.rescale(x_in.next(), (0, 1), (-1, 1))
        back = hem.to_nchw(y)
        torch.cuda.synchronize()
        assert y.shape == (B, H, H, C) and back.shape == (B, C, H, H)
        assert torch.equal(back.torch().permute(0, 2, 3, 1).contiguous(), y.torch().float())
        outs.append(y.torch().float().cpu())
    assert float((outs[0] - outs[1]).abs().max()) < 1e-2
    want = (raw.float() / 255 * 2 - 1).permute(0, 2, 3, 1)
    assert float((outs[0] - want).abs().max()) < 1e-2


def test_summaries_match_torch_reductions():
    """tensor_stats = histogram moments + zero fraction in one pass; montage layout of ops/summaries.py:113-117."""
    from b200gan.ops import summaries as SM
    E.begin()
    gen = torch.Generator().manual_seed(6)
    v = torch.randn(300000, generator=gen) * 3
    v[::7] = 0
    for t in (v.cuda(), v.cuda().to(torch.bfloat16)):
        st = SM.tensor_stats(t)
        torch.cuda.synchronize()
        f = t.float().cpu()
        h = st["histogram"]
        assert abs(h["min"] - float(f.min())) < 1e-6 and abs(h["max"] - float(f.max())) < 1e-6
        assert abs(h["sum"] - float(f.double().sum())) < 1e-2 * float(f.abs().double().sum()) ** 0.5 + 1.0
        assert abs(h["sum_squares"] - float((f.double() ** 2).sum())) < 1e-3 * float((f.double() ** 2).sum())
        assert abs(st["sparsity"] - float((f == 0).float().mean())) < 1e-6
        counts = h["bucket_counts"].long()
        assert int(counts.sum()) == f.numel()
        half = SM.N_BUCKETS // 2
        assert int(counts[:half].sum()) == int((f < 0).sum())                       # negatives on the left half
        assert int(counts[half]) == int(((f >= 0) & (f.abs() < 1e-12)).sum())        # the bucket around zero
    imgs = torch.arange(6 * 2 * 3 * 1, dtype=torch.float32).reshape(6, 2, 3, 1).cuda()
    mont = SM.montage_summary(imgs, 2, 3)
    torch.cuda.synchronize()
    # reference: split the batch into n=3 groups of m=2, stack each group vertically, groups side by side
    ref = torch.cat([torch.cat([imgs[g * 2 + r] for r in range(2)], dim=0) for g in range(3)], dim=1)
    assert torch.equal(mont, ref)
    assert SM.factorization(64) == (8, 8) and SM.factorization(12) == (3, 4)


def test_summary_pass_collects_layer_outputs():
    """A forward pass under summaries.collecting(store) registers every non-reused layer output like the reference's
    'conv_layers' / 'dense_layers' collections; summarize_* return histogram / sparsity / montage per tensor."""
    from b200gan.models import gan as gan_model
    from b200gan.ops import summaries as SM
    args = argparse.Namespace(model="iwgan", batch_size=8, latent_size=16, n_disc_train=1, optimizer="adam", lr=1e-4,
                              beta1=0.5, beta2=0.9)
    sess = S.Session(seed=0)
    sess.use_graphs = False
    x_in = S.Input(8, (32, 32, 3), slots=2)
    train = gan_model.gan(x_in, args)
    x_in.ring.copy_(torch.rand(x_in.ring.shape, device="cuda"))
    sess.begin_step(); x_in.reset()
    with SM.collecting(sess.store):
        gl, dl = train.tower(x_in.next(), 'g')
        acts = SM.summarize_activations(sess.store)
    torch.cuda.synchronize()
    # generator fc1 + dc1..dc3, critic c1..c3 + fc2 (the reuse=True critic passes add nothing: ops/layers.py:60)
    assert sorted(k.split("/")[-1] for k in acts) == sorted(["fc1", "dc1", "dc2", "dc3", "c1", "c2", "c3", "fc2"])
    c1 = acts["activations/discriminator/c1"]
    assert c1["montage"].shape == (10 * 16, 20 * 16, 1) or c1["montage"].shape[2] == 1     # 200 channel maps of 16x16
    assert 0.0 <= c1["sparsity"] < 0.05
    w = SM.summarize_weights_biases(sess.store)
    assert "weights/discriminator/vars/c2/weights" in w and w["weights/discriminator/vars/c2/weights"]["histogram"]["num"] == 5 * 5 * 16 * 32
    assert set(SM.summarize_losses({"g_loss": gl.buf, "d_loss": dl.buf})) == {"loss/g_loss", "loss/d_loss"}
