"""Reference model graphs restated on torch-CPU autograd (test infrastructure; PARITY UNPINNED).

Follows /root/reference/models/gan.py, models/cnn.py, models/vae.py, util.py with the
minimal shape generalisation of SURVEY.md Appendix C #1 (at 64x64x3 it is the reference,
layer for layer).  Parameters are plain dicts keyed by the TF variable names the reference
would create (`generator/vars/dc1/weights`, `generator/BatchNorm_1/beta`, ...), noise is
always an explicit input (A.8), and gradients come from torch autograd over the restated ops
(an independent derivation from the hand-scheduled backward of the CUDA path).
"""
import math
from collections import OrderedDict

import torch

from . import tf_ops as T


# =========================================================================== shapes
def n_up_stages(H):
    """Generator deconv count so that 4 * 2^n == H (models/gan.py:247-252 has n=4 at H=64)."""
    n = int(round(math.log2(H / 4)))
    assert 4 * 2 ** n == H, "GAN family needs H = 4*2^n"
    return n


def gan_param_specs(model, H, C, L):
    """Ordered {tf_name: shape} in TF creation order (models/gan.py:234-287; ops/layers.py).
    BN betas: generator fc1,dc1..dc(n-1); discriminator c2,c3 for gan/wgan, twice (real and
    fake paths do not share them)."""
    n = n_up_stages(H)
    g = OrderedDict()
    bn = 0

    def bn_name(scope, k):
        return "%s/BatchNorm%s/beta" % (scope, "" if k == 0 else "_%d" % k)

    g["generator/vars/fc1/weights"] = (L, 64 * L)
    g["generator/vars/fc1/bias"] = (64 * L,)
    g[bn_name("generator", bn)] = (64 * L,); bn += 1
    cin = 4 * L
    for i in range(1, n + 1):
        last = i == n
        cout = C if last else cin // 2
        g["generator/vars/dc%d/weights" % i] = (5, 5, cout, cin)
        g["generator/vars/dc%d/bias" % i] = (cout,)
        if not last:
            g[bn_name("generator", bn)] = (cout,); bn += 1
        cin = cout
    d = OrderedDict()
    d["discriminator/vars/c1/weights"] = (5, 5, C, L)
    d["discriminator/vars/c1/bias"] = (L,)
    d["discriminator/vars/c2/weights"] = (5, 5, L, 2 * L)
    d["discriminator/vars/c2/bias"] = (2 * L,)
    if model != "iwgan":
        d[bn_name("discriminator", 0)] = (2 * L,)
    d["discriminator/vars/c3/weights"] = (5, 5, 2 * L, 4 * L)
    d["discriminator/vars/c3/bias"] = (4 * L,)
    if model != "iwgan":
        d[bn_name("discriminator", 1)] = (4 * L,)
    d["discriminator/vars/fc2/weights"] = (64 * L, 1)
    d["discriminator/vars/fc2/bias"] = (1,)
    if model != "iwgan":
        d[bn_name("discriminator", 2)] = (2 * L,)
        d[bn_name("discriminator", 3)] = (4 * L,)
    return g, d


def ae_param_specs(model, H, C, L):
    """cnn / vae variables (models/cnn.py:82-134, models/vae.py:93-151), generalised:
    bottleneck s = H after four k5s2 convs (4 at H=64)."""
    sizes = [H]
    for _ in range(4):
        sizes.append(-(-sizes[-1] // 2))
    s = sizes[-1]
    p = OrderedDict()
    bn = 0
    chans = [(C, 64, 5), (64, 128, 5), (128, 256, 5), (256, 256, 5), (256, 96, 1), (96, 32, 1)]
    for i, (ci, co, k) in enumerate(chans, 1):
        p["encoder/vars/c%d/weights" % i] = (k, k, ci, co)
        p["encoder/vars/c%d/bias" % i] = (co,)
        if model == "vae":
            p["encoder/BatchNorm%s/beta" % ("" if bn == 0 else "_%d" % bn)] = (co,); bn += 1
    p["latent/vars/d1/weights"] = (32 * s * s, L)
    p["latent/vars/d1/bias"] = (L,)
    if model == "vae":
        p["latent/vars/d2/weights"] = (32 * s * s, L)
        p["latent/vars/d2/bias"] = (L,)
    p["decoder/vars/d1/weights"] = (L, 32 * s * s)
    p["decoder/vars/d1/bias"] = (32 * s * s,)
    p["decoder/vars/c1/weights"] = (1, 1, 32, 96)
    p["decoder/vars/c1/bias"] = (96,)
    p["decoder/vars/c2/weights"] = (1, 1, 96, 256)
    p["decoder/vars/c2/bias"] = (256,)
    for i, (ci, co) in enumerate([(256, 256), (256, 128), (128, 64), (64, C)], 1):
        p["decoder/vars/dc%d/weights" % i] = (5, 5, co, ci)
        p["decoder/vars/dc%d/bias" % i] = (co,)
    return p, sizes


def init_params(specs, seed=0, dtype=torch.float32):
    """Xavier-uniform for weights AND biases, zeros for BN beta (A.3, A.7)."""
    gen = torch.Generator().manual_seed(seed)
    out = OrderedDict()
    for name, shape in specs.items():
        if name.endswith("/beta"):
            out[name] = torch.zeros(shape, dtype=dtype)
        else:
            out[name] = T.xavier_uniform(tuple(shape), gen, dtype)
    return out


# =========================================================================== GAN family
def generator(p, z, H, C, L):
    """models/gan.py:234-254."""
    n = n_up_stages(H)
    bn = 0

    def beta():
        nonlocal bn
        name = "generator/BatchNorm%s/beta" % ("" if bn == 0 else "_%d" % bn)
        bn += 1
        return p[name]

    y = T.dense(z, p["generator/vars/fc1/weights"], p["generator/vars/fc1/bias"], beta(), "relu")
    y = y.reshape(-1, 4, 4, 4 * L)
    for i in range(1, n + 1):
        last = i == n
        y = T.deconv2d(y, p["generator/vars/dc%d/weights" % i], p["generator/vars/dc%d/bias" % i], 2,
                       None if last else beta(), "tanh" if last else "relu")
    return y.reshape(-1, H * H * C)


def discriminator(p, x, H, C, L, model, bn_base=0):
    """models/gan.py:257-287.  bn_base selects which BatchNorm betas this call created
    (0 for D(real), 2 for D(fake) in gan/wgan; unused for iwgan)."""
    use_bn = model != "iwgan"
    final = "sigmoid" if model == "gan" else None

    def beta(k):
        if not use_bn:
            return None
        k += bn_base
        return p["discriminator/BatchNorm%s/beta" % ("" if k == 0 else "_%d" % k)]

    x = x.reshape(-1, H, H, C)
    x = T.conv2d(x, p["discriminator/vars/c1/weights"], p["discriminator/vars/c1/bias"], 2, None, "lrelu")
    x = T.conv2d(x, p["discriminator/vars/c2/weights"], p["discriminator/vars/c2/bias"], 2, beta(0), "lrelu")
    x = T.conv2d(x, p["discriminator/vars/c3/weights"], p["discriminator/vars/c3/bias"], 2, beta(1), "lrelu")
    x = x.reshape(-1, 4 * 4 * 4 * L)          # 4 rows per image at H=64 (App. C #2)
    x = T.dense(x, p["discriminator/vars/fc2/weights"], p["discriminator/vars/fc2/bias"], None, final)
    return x.reshape(-1)


def gan_losses(p, x01, z, alpha, model, H, C, L):
    """models/gan.py:49-50 (rescale), 55-63, 178-231.  x01: [B,H,W,C] in [0,1].
    Returns (g_loss, d_loss, g)."""
    x = T.stored(2 * (x01.reshape(x01.shape[0], -1) - 0.5))
    g = generator(p, z, H, C, L)
    d_real = discriminator(p, x, H, C, L, model, 0)
    d_fake = discriminator(p, g, H, C, L, model, 2)
    if model == "gan":
        g_loss = torch.mean(-torch.log(d_fake + 1e-8))
        d_loss = torch.mean(-torch.log(d_real + 1e-8) - torch.log(1 - d_fake + 1e-8))
    elif model == "wgan":
        g_loss = -d_fake.mean()
        d_loss = d_fake.mean() - d_real.mean()
    else:
        g_loss = -d_fake.mean()
        interp = T.stored(x + alpha * (g - x))
        if not interp.requires_grad:                     # pure evaluation (no trainable inputs)
            interp = interp.detach().requires_grad_(True)
        d_int = discriminator(p, interp, H, C, L, model)
        grads = torch.autograd.grad(d_int.sum(), interp, create_graph=True)[0]
        slopes = torch.sqrt(torch.sum(grads ** 2))        # ONE norm over the tower batch (App. C #3)
        gp = (slopes - 1.0) ** 2
        d_loss = d_fake.mean() - d_real.mean() + 10.0 * gp
    return g_loss, d_loss, g


def iwgan_critic_grad_terms(p, x01, z, alpha, H, C, L):
    """The three additive pieces of the IWGAN critic gradient (fake, real, penalty) separately:
    parity tests scale their tolerance by the sum of the pieces' norms because the pieces cancel."""
    names_d = [k for k in p if k.startswith("discriminator/")]
    q = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in p.items())
    x = T.stored(2 * (x01.reshape(x01.shape[0], -1) - 0.5))
    g = generator(q, z, H, C, L).detach()
    d_real = discriminator(q, x, H, C, L, "iwgan")
    d_fake = discriminator(q, g, H, C, L, "iwgan")
    interp = T.stored(x + alpha * (g - x)).detach().requires_grad_(True)
    d_int = discriminator(q, interp, H, C, L, "iwgan")
    grads = torch.autograd.grad(d_int.sum(), interp, create_graph=True)[0]
    gp = 10.0 * (torch.sqrt(torch.sum(grads ** 2)) - 1.0) ** 2
    terms = []
    for loss in (d_fake.mean(), -d_real.mean(), gp):
        gs = torch.autograd.grad(loss, [q[k] for k in names_d], retain_graph=True, allow_unused=True)
        terms.append(OrderedDict((k, torch.zeros_like(q[k]) if v is None else v) for k, v in zip(names_d, gs)))
    return terms


def gan_grads(p, x01, z, alpha, model, H, C, L):
    """compute_gradients(g_loss, g_params) and (d_loss, d_params) — models/gan.py:65-68."""
    names_g = [k for k in p if k.startswith("generator/")]
    names_d = [k for k in p if k.startswith("discriminator/")]
    q = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in p.items())
    g_loss, d_loss, g = gan_losses(q, x01, z, alpha, model, H, C, L)
    gg = torch.autograd.grad(g_loss, [q[k] for k in names_g], retain_graph=True, allow_unused=True)
    dg = torch.autograd.grad(d_loss, [q[k] for k in names_d], allow_unused=True)
    grads = OrderedDict()
    for k, v in zip(names_g, gg):
        grads[k] = torch.zeros_like(q[k]) if v is None else v
    for k, v in zip(names_d, dg):
        grads[k] = torch.zeros_like(q[k]) if v is None else v
    return {"g_loss": g_loss.detach(), "d_loss": d_loss.detach(), "g": g.detach(), "grads": grads}


# =========================================================================== cnn / vae
def ae_encoder(p, x, model):
    bn = 0
    for i, s in enumerate([2, 2, 2, 2, 1, 1], 1):
        beta = None
        if model == "vae":
            beta = p["encoder/BatchNorm%s/beta" % ("" if bn == 0 else "_%d" % bn)]
            bn += 1
        x = T.conv2d(x, p["encoder/vars/c%d/weights" % i], p["encoder/vars/c%d/bias" % i], s, beta, "lrelu")
    return x


def ae_decoder(p, z, sizes, final_act):
    s = sizes[-1]
    x = T.dense(z, p["decoder/vars/d1/weights"], p["decoder/vars/d1/bias"], None, "relu")
    x = x.reshape(-1, s, s, 32)
    x = T.conv2d(x, p["decoder/vars/c1/weights"], p["decoder/vars/c1/bias"], 1, None, "relu")
    x = T.conv2d(x, p["decoder/vars/c2/weights"], p["decoder/vars/c2/bias"], 1, None, "relu")
    for i in range(1, 5):
        out = sizes[4 - i]
        x = T.deconv2d(x, p["decoder/vars/dc%d/weights" % i], p["decoder/vars/dc%d/bias" % i], 2, None,
                       final_act if i == 4 else "relu", out_hw=(out, out), store_out=(i != 4))
    return x


def cnn_losses(p, x01, sizes):
    """models/cnn.py:20-79: x -> 2(x-0.5); loss = mean|x - d|."""
    x = T.stored(2 * (x01 - 0.5))
    e = ae_encoder(p, x, "cnn")
    z = T.dense(e.reshape(e.shape[0], -1), p["latent/vars/d1/weights"], p["latent/vars/d1/bias"])
    d = ae_decoder(p, z, sizes, "tanh")
    return {"loss": T.l1_mean(d, x)}, d


def vae_losses(p, x01, eps, sizes):
    """models/vae.py:25-129: x stays in [0,1]; z = mu + sigma*eps; Bernoulli recon (sum) + KL;
    only decoder_loss is differentiated (models/vae.py:41)."""
    e = ae_encoder(p, x01, "vae")
    flat = e.reshape(e.shape[0], -1)
    mu = T.dense(flat, p["latent/vars/d1/weights"], p["latent/vars/d1/bias"])
    sd = T.dense(flat, p["latent/vars/d2/weights"], p["latent/vars/d2/bias"])
    z = T.stored(mu + sd * eps)
    d = ae_decoder(p, z, sizes, "sigmoid")
    rec = -torch.sum(x01 * torch.log(1e-8 + d) + (1 - x01) * torch.log(1e-8 + (1 - d)))
    kl = 0.5 * torch.sum(mu ** 2 + sd ** 2 - torch.log(1e-8 + sd ** 2) - 1)
    return {"decoder_loss": rec, "latent_loss": kl, "total_loss": rec + kl}, d


def ae_grads(p, x01, eps, model, sizes):
    q = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in p.items())
    if model == "cnn":
        losses, d = cnn_losses(q, x01, sizes)
        obj = losses["loss"]
    else:
        losses, d = vae_losses(q, x01, eps, sizes)
        obj = losses["decoder_loss"]
    gs = torch.autograd.grad(obj, list(q.values()), allow_unused=True)
    grads = OrderedDict((k, torch.zeros_like(q[k]) if g is None else g) for k, g in zip(q, gs))
    return {"losses": {k: v.detach() for k, v in losses.items()}, "out": d.detach(), "grads": grads}


# =========================================================================== training schedules
class AdamState:
    """One tf.train.AdamOptimizer instance (util.py:178-181): own step counter and slots."""

    def __init__(self, params, names, lr, beta1, beta2):
        self.names, self.lr, self.b1, self.b2, self.t = list(names), lr, beta1, beta2, 0
        self.m = {k: torch.zeros_like(params[k]) for k in self.names}
        self.v = {k: torch.zeros_like(params[k]) for k in self.names}

    def apply(self, params, grads):
        self.t += 1
        for k in self.names:
            T.adam_step(params[k], grads[k], self.m[k], self.v[k], self.t, self.lr, self.b1, self.b2)


class GanTrainer:
    """models/gan.py:110-175 schedules on one tower.  `noise(kind, shape)` supplies z / alpha and
    `batches()` yields a fresh [B,H,W,C] batch per sess.run-equivalent (App. C #8)."""

    def __init__(self, model, H, C, L, B, lr=1e-4, beta1=0.5, beta2=0.9, n_disc=5, seed=0,
                 dtype=torch.float32):
        self.model, self.H, self.C, self.L, self.B, self.n_disc = model, H, C, L, B, n_disc
        gs, ds = gan_param_specs(model, H, C, L)
        specs = OrderedDict(list(gs.items()) + list(ds.items()))
        self.p = init_params(specs, seed, dtype)
        self.g_names, self.d_names = list(gs), list(ds)
        self.g_opt = AdamState(self.p, self.g_names, lr, beta1, beta2)
        self.d_opt = AdamState(self.p, self.d_names, lr, beta1, beta2)

    def _run(self, x01, z, alpha):
        return gan_grads(self.p, x01, z, alpha, self.model, self.H, self.C, self.L)

    def _clip(self, names):
        for k in names:                                   # models/gan.py:142-143
            self.p[k].clamp_(-0.01, 0.01)

    def iteration(self, next_batch, next_noise):
        """One train_func call (train.py:307).  next_noise() -> (z, alpha)."""
        if self.model == "gan":                           # one run, both updates (App. C #7)
            r = self._run(next_batch(), *next_noise())
            self.d_opt.apply(self.p, r["grads"])
            self.g_opt.apply(self.p, r["grads"])
            return {"g_loss": float(r["g_loss"]), "d_loss": float(r["d_loss"])}
        for _ in range(self.n_disc):
            r = self._run(next_batch(), *next_noise())
            if self.model == "wgan":
                self._clip(self.d_names)
            self.d_opt.apply(self.p, r["grads"])
        r = self._run(next_batch(), *next_noise())
        if self.model == "wgan":
            self._clip(self.g_names)
        self.g_opt.apply(self.p, r["grads"])
        return {"g_loss": float(r["g_loss"]), "d_loss": float(r["d_loss"])}
