"""pix2pix restated on torch-CPU (test infrastructure; PARITY UNPINNED) — follows
/root/reference/hem/models/pix2pix.py:81-304 and hem/ops/layers.py:71-211 (NCHW there, NHWC here: the
axis-1 concatenations are last-axis concatenations)."""
from collections import OrderedDict

import torch

from . import tf_ops as T

ENC = [(3, 64), (64, 128), (128, 256), (256, 512), (512, 512), (512, 512), (512, 512), (512, 512)]
DEC = [(512, 512), (1024, 512), (1024, 512), (1024, 512), (1024, 256), (512, 128), (256, 64), (128, 1)]
DISC = [(4, 64), (64, 128), (128, 256), (256, 512), (512, 1)]


def param_specs(batch_norm_gen=False, batch_norm_disc=False):
    """TF creation order/names: generator/enocder/vars/<i>/..., generator/decoder/vars/<i>/... with the
    decoder's BatchNorm[_k]/beta interleaved, discriminator/vars/m<i>/... (pix2pix.py:182-256)."""
    g = OrderedDict()
    bn = 0
    for i, (ci, co) in enumerate(ENC, 1):
        g["generator/enocder/vars/%d/weights" % i] = (4, 4, ci, co)
        g["generator/enocder/vars/%d/bias" % i] = (co,)
        if batch_norm_gen and i > 1:
            g["generator/enocder/BatchNorm%s/beta" % ("" if bn == 0 else "_%d" % bn)] = (co,); bn += 1
    bn = 0
    for i, (ci, co) in enumerate(DEC, 1):
        g["generator/decoder/vars/%d/weights" % i] = (4, 4, co, ci)
        g["generator/decoder/vars/%d/bias" % i] = (co,)
        g["generator/decoder/BatchNorm%s/beta" % ("" if bn == 0 else "_%d" % bn)] = (co,); bn += 1
    d = OrderedDict()
    for i, (ci, co) in enumerate(DISC, 1):
        d["discriminator/vars/m%d/weights" % i] = (4, 4, ci, co)
        d["discriminator/vars/m%d/bias" % i] = (co,)
    if batch_norm_disc:
        raise NotImplementedError("oracle: --batch_norm_disc variant not restated")
    return g, d


def init_params(specs, seed=0, dtype=torch.float32):
    """random_normal_initializer(0, 0.02) for weights AND biases (pix2pix.py:180,201,248); beta = 0."""
    gen = torch.Generator().manual_seed(seed)
    out = OrderedDict()
    for name, shape in specs.items():
        if name.endswith("/beta"):
            out[name] = torch.zeros(shape, dtype=dtype)
        else:
            out[name] = (torch.randn(shape, generator=gen, dtype=torch.float64) * 0.02).to(dtype)
    return out


def generator(p, x):
    es = []
    h = x
    for i in range(1, 9):
        h = T.conv2d(h, p["generator/enocder/vars/%d/weights" % i], p["generator/enocder/vars/%d/bias" % i], 2, None,
                     "lrelu")
        es.append(h)
    y = es[7]
    for i in range(1, 9):
        beta = p["generator/decoder/BatchNorm%s/beta" % ("" if i == 1 else "_%d" % (i - 1))]
        if i > 1:
            y = torch.cat([y, es[8 - i]], dim=-1)
        y = T.deconv2d(y, p["generator/decoder/vars/%d/weights" % i], p["generator/decoder/vars/%d/bias" % i], 2, beta,
                       "tanh" if i == 8 else "relu")         # hem.lrelu(x, leak=0) == relu
    return y


def discriminator(p, x, y):
    h = torch.cat([x, y], dim=-1)
    for i in range(1, 6):
        h = T.conv2d(h, p["discriminator/vars/m%d/weights" % i], p["discriminator/vars/m%d/bias" % i], 2, None,
                     None if i == 5 else "lrelu")
    return h


def losses(p, x01, y01, add_l1=False):
    x = T.stored(2 * x01 - 1)
    y = T.stored(2 * y01 - 1)
    g = generator(p, x)
    dr = discriminator(p, x, y)
    df = discriminator(p, x, g)
    g01 = T.stored((g + 1) * 0.5)
    yy = T.stored((y + 1) * 0.5)
    g_fake = T.sigmoid_ce(df, torch.ones_like(df)).mean()
    l1 = T.l1_mean(g01, yy)
    g_total = g_fake + 10.0 * l1 if add_l1 else g_fake
    d_real = T.sigmoid_ce(dr, torch.ones_like(dr)).mean()
    d_fake = T.sigmoid_ce(df, torch.zeros_like(df)).mean()
    return {"l1": l1, "g_fake": g_fake, "g_total": g_total, "d_real": d_real, "d_fake": d_fake,
            "d_total": d_real + d_fake, "rmse": T.rmse(yy, g01)}


def grads(p, x01, y01, add_l1=False):
    q = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in p.items())
    ls = losses(q, x01, y01, add_l1)
    gn = [k for k in q if k.startswith("generator/")]
    dn = [k for k in q if k.startswith("discriminator/")]
    gg = torch.autograd.grad(ls["g_total"], [q[k] for k in gn], retain_graph=True, allow_unused=True)
    dg = torch.autograd.grad(ls["d_total"], [q[k] for k in dn], allow_unused=True)
    out = OrderedDict()
    for k, v in list(zip(gn, gg)) + list(zip(dn, dg)):
        out[k] = torch.zeros_like(q[k]) if v is None else v
    return {"losses": {k: float(v) for k, v in ls.items()}, "grads": out}
