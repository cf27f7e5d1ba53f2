"""TensorFlow-1.x op semantics restated on torch-CPU (test infrastructure; PARITY UNPINNED).

Each function cites the reference call site it stands in for
(paths relative to /root/reference) and the TF behaviour it restates
(SURVEY.md Appendix A).  All tensors are NHWC unless stated, dtype follows the
input (float32 for timing/parity, float64 for finite-difference checks).
"""
import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- storage-precision emulation
class _RoundBF16(torch.autograd.Function):
    """Round to bf16 in the forward, identity in the backward (straight-through)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        # store_bf16(grads=True): the gradient arriving at a stored tensor is itself a tensor the CUDA path
        # keeps in bf16, so it is rounded as well (used to measure how much of a GPU-vs-fp32-oracle difference
        # is storage rounding: tests/test_oracle.py::test_bf16_storage_noise_*)
        return g.to(torch.bfloat16).to(g.dtype) if _ROUND_GRADS else g


_STORE_BF16 = False
_ROUND_GRADS = False


class store_bf16:
    """Context: emulate the CUDA path's storage precision — every tensor it writes to HBM as bf16
    (conv/dense outputs before batch-norm, layer outputs after the activation, the rescaled input, the
    GP interpolates) is rounded to bf16 here too.  Arithmetic stays fp32.  With it on, ReLU/LReLU masks
    agree with the GPU's, which removes the mask-flip noise that dominates fp32-oracle comparisons."""

    def __init__(self, on=True, grads=False):
        self.on, self.grads = on, grads

    def __enter__(self):
        global _STORE_BF16, _ROUND_GRADS
        self.prev, _STORE_BF16, _ROUND_GRADS = (_STORE_BF16, _ROUND_GRADS), self.on, self.on and self.grads

    def __exit__(self, *a):
        global _STORE_BF16, _ROUND_GRADS
        _STORE_BF16, _ROUND_GRADS = self.prev


def stored(t):
    return _RoundBF16.apply(t) if _STORE_BF16 else t


# --------------------------------------------------------------------------- decision injection
class inject_decisions:
    """Context: the oracle takes every DISCONTINUOUS decision of the graph — the relu / lrelu masks
    (ops/activations.py:28) and the sign of the L1 residual (models/cnn.py:77) — from a
    queue recorded by the implementation under test instead of from its own pre-activations; all
    arithmetic stays plain fp32 (no storage emulation).

    Why: d relu/dx jumps at 0, so a pre-activation that differs by one bf16 ulp between the two sides
    flips the unit and changes its gradient by O(1); that noise (a few % per ReLU layer) hides real
    errors of the same size.  With the decisions pinned, the two sides evaluate the SAME piecewise-
    linear function and gradients must agree to bf16 rounding.  The queue is audited while it is
    consumed: `stats` records, per decision, the fraction of units whose injected decision differs from
    the oracle's own and how large the oracle's pre-activation is at those units relative to the layer
    rms — an implementation whose masks are wrong (not merely rounded differently) shows up there.
    """

    def __init__(self, decisions):
        self.queue = list(decisions)
        self.stats = []

    def __enter__(self):
        global _INJECT
        self.prev, _INJECT = _INJECT, self
        return self

    def __exit__(self, *a):
        global _INJECT
        _INJECT = self.prev

    def pop(self, kind, like):
        """Next recorded decision as a float tensor shaped like `like`: +1 / 0 for an activation mask
        (unit passes / is cut), +1 / 0 / -1 for an L1 residual sign.  `like` is the oracle's own
        pre-activation (residual); the audit compares its sign with the recorded decision."""
        if not self.queue:
            raise AssertionError("decision queue exhausted at a %s of shape %s" % (kind, tuple(like.shape)))
        k, m = self.queue.pop(0)
        if k != kind or m.numel() != like.numel():
            raise AssertionError("decision queue out of step: got %s %s, oracle is at %s %s"
                                 % (k, tuple(m.shape), kind, tuple(like.shape)))
        m = m.reshape(like.shape).to(like.dtype)
        v = like.detach()
        own = (v > 0).to(like.dtype) if kind == "act" else torch.sign(v)
        flip = m != own
        rms = float(v.pow(2).mean().sqrt())
        at = float(v[flip].abs().mean()) if bool(flip.any()) else 0.0
        self.stats.append({"kind": kind, "shape": tuple(like.shape), "flip_frac": float(flip.float().mean()),
                           "flip_mag_over_rms": at / max(rms, 1e-30)})
        return m


_INJECT = None


class record_decisions(inject_decisions):
    """Context: run the oracle on its own decisions and append them to `self.queue` in the format
    `inject_decisions` consumes (used to test the injection machinery against itself)."""

    def __init__(self):
        inject_decisions.__init__(self, [])

    def pop(self, kind, like):
        v = like.detach()
        m = (v > 0).to(like.dtype) if kind == "act" else torch.sign(v)
        self.queue.append((kind, m.clone()))
        return m


def l1_mean(a, b):
    """mean(|a - b|) — models/cnn.py:77, hem/models/pix2pix.py:286.  Under `inject_decisions` the sign
    of the residual is the recorded one (the loss stays piecewise linear in `a`)."""
    r = a - b
    if _INJECT is None:
        return torch.mean(torch.abs(r))
    return torch.mean(r * _INJECT.pop("l1", r))


# --------------------------------------------------------------------------- padding
def same_pad(in_size, k, s):
    """TF 'SAME' padding (A.1): out = ceil(in/s); pad_total = max((out-1)s+k-in, 0);
    before = total//2 (the smaller half goes FIRST).  Used by every tf.nn.conv2d /
    conv2d_transpose call in ops/layers.py:101,142 and hem/ops/layers.py:118,189."""
    out = -(-in_size // s)
    total = max((out - 1) * s + k - in_size, 0)
    before = total // 2
    return out, before, total - before


# --------------------------------------------------------------------------- conv
def conv2d_same(x, K, stride):
    """tf.nn.conv2d(x, K, [1,s,s,1], 'SAME') — ops/layers.py:101.
    x [N,H,W,Cin], K [kh,kw,Cin,Cout] (cross-correlation, no flip)."""
    kh, kw = K.shape[0], K.shape[1]
    _, pt, pb = same_pad(x.shape[1], kh, stride)
    _, pl, pr = same_pad(x.shape[2], kw, stride)
    xn = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb))
    y = F.conv2d(xn, K.permute(3, 2, 0, 1), stride=stride)
    return y.permute(0, 2, 3, 1)


def conv2d_transpose_same(x, K, out_hw, stride):
    """tf.nn.conv2d_transpose(x, K, output_shape, [1,s,s,1], 'SAME') — ops/layers.py:142.
    x [N,h,w,Cin_t], K [kh,kw,Cout_t,Cin_t]; exact adjoint of conv2d_same mapping
    [N,Hout,Wout,Cout_t] -> x.shape (A.2): full transposed conv then crop at the
    forward conv's pad_before."""
    kh, kw = K.shape[0], K.shape[1]
    Hout, Wout = out_hw
    _, pt, _ = same_pad(Hout, kh, stride)
    _, pl, _ = same_pad(Wout, kw, stride)
    # underlying conv weight is [Cout_conv=Cin_t, Cin_conv=Cout_t, kh, kw]; conv_transpose2d
    # takes weight [in_channels=Cin_t, out_channels=Cout_t, kh, kw]
    full = F.conv_transpose2d(x.permute(0, 3, 1, 2), K.permute(3, 2, 0, 1), stride=stride)
    # the forward conv may not touch the last rows of its padded input (e.g. k5 s2 on an even
    # size): the adjoint is zero there, so zero-extend before cropping.
    need_h, need_w = pt + Hout, pl + Wout
    if full.shape[2] < need_h or full.shape[3] < need_w:
        full = F.pad(full, (0, max(need_w - full.shape[3], 0), 0, max(need_h - full.shape[2], 0)))
    y = full[:, :, pt:pt + Hout, pl:pl + Wout]
    return y.permute(0, 2, 3, 1)


def conv2d_valid(x, K, stride):
    """tf.nn.conv2d(x, K, strides, 'VALID') — hem/ops/layers.py:118 with padding='VALID'."""
    y = F.conv2d(x.permute(0, 3, 1, 2), K.permute(3, 2, 0, 1), stride=stride)
    return y.permute(0, 2, 3, 1)


def instance_norm(x, scale, shift, eps=1e-3):
    """hem/ops/images.py:73-89 (NHWC here): per sample and channel moments over H, W; scale * xhat + shift."""
    mu = x.mean(dim=(1, 2), keepdim=True)
    var = ((x - mu) ** 2).mean(dim=(1, 2), keepdim=True)
    return scale * (x - mu) / torch.sqrt(var + eps) + shift


def dropout(x, keep_prob, u):
    """tf.nn.dropout(x, keep_prob) with its uniform draw u made explicit: x / keep * floor(keep + u)."""
    return x / keep_prob * torch.floor(keep_prob + u)


# --------------------------------------------------------------------------- activations
def lrelu(x, leak=0.2):
    """ops/activations.py:28 — tf.maximum(leak*x, x).  TF's Maximum gradient routes to the
    first argument where leak*x >= x, i.e. slope = leak for x <= 0 (A.5)."""
    if _INJECT is not None:
        pos = _INJECT.pop("act", x)
        return x * (pos + (1 - pos) * leak)
    return _LReLU.apply(x, leak)


def relu(x):
    if _INJECT is not None:
        return x * _INJECT.pop("act", x)
    return torch.relu(x)


class _LReLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, leak):
        ctx.save_for_backward(x)
        ctx.leak = leak
        return torch.maximum(leak * x, x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        slope = torch.where(x > 0, torch.ones_like(x), torch.full_like(x, ctx.leak))
        return g * slope, None


ACTIVATIONS = {
    None: lambda t: t,
    "none": lambda t: t,
    "relu": relu,
    "lrelu": lrelu,
    "tanh": torch.tanh,
    "sigmoid": torch.sigmoid,
}


# --------------------------------------------------------------------------- batch norm
def batch_norm_train(h, beta, eps=1e-3):
    """tf.contrib.layers.batch_norm(h) with defaults — ops/layers.py:58,103,144 (A.3):
    training mode always, scale=False (no gamma), center=True, eps 1e-3, biased batch
    variance over every axis but the last (channels)."""
    axes = tuple(range(h.dim() - 1))
    mean = h.mean(dim=axes, keepdim=True)
    var = ((h - mean) ** 2).mean(dim=axes, keepdim=True)
    return (h - mean) * torch.rsqrt(var + eps) + beta


def batch_norm_moving_update(moving_mean, moving_var, h, decay=0.999, unbiased=False):
    """UPDATE_OPS side effect of batch_norm (A.3): mv -= (mv - batch)*(1-decay)."""
    axes = tuple(range(h.dim() - 1))
    n = h.numel() // h.shape[-1]
    mean = h.mean(dim=axes)
    var = ((h - mean) ** 2).mean(dim=axes)
    if unbiased:
        var = var * n / max(n - 1, 1)
    return (moving_mean - (moving_mean - mean) * (1 - decay),
            moving_var - (moving_var - var) * (1 - decay))


# --------------------------------------------------------------------------- layers (ops/layers.py)
def dense(x, W, b, beta=None, activation=None):
    """ops/layers.py:27-62: act(BN(xW + b))."""
    h = x @ W + b
    if beta is not None:
        h = batch_norm_train(stored(h), beta)
    out = ACTIVATIONS[activation](h)
    return out if (W.shape[1] == 1) else stored(out)       # 1-unit dense keeps fp32 (GEMV kernel)


def conv2d(x, K, b, stride, beta=None, activation=None):
    """ops/layers.py:66-107: act(BN(conv_SAME(x,K) + b))."""
    h = conv2d_same(x, K, stride) + b
    if beta is not None:
        h = batch_norm_train(stored(h), beta)
    return stored(ACTIVATIONS[activation](h))


def deconv2d(x, K, b, stride=2, beta=None, activation=None, out_hw=None, store_out=True):
    """ops/layers.py:111-148: act(BN(conv2d_transpose_SAME(x,K) + b)); output is 2x the input
    (ops/layers.py:141) unless out_hw is given (hem/ops/layers.py:185-187).  store_out=False: the CUDA
    path keeps this output in fp32 (the autoencoders' loss heads), so storage emulation leaves it alone."""
    if out_hw is None:
        out_hw = (x.shape[1] * 2, x.shape[2] * 2)
    h = conv2d_transpose_same(x, K, out_hw, stride) + b
    if beta is not None:
        h = batch_norm_train(stored(h), beta)
    out = ACTIVATIONS[activation](h)
    return stored(out) if store_out else out


# --------------------------------------------------------------------------- initialisers
def xavier_uniform(shape, gen, dtype=torch.float32):
    """tf.contrib.layers.xavier_initializer() (A.7): U(+-sqrt(6/(fan_in+fan_out))).
    conv [k,k,i,o]: fan_in=k*k*i, fan_out=k*k*o; dense [i,o]; bias [o]: fan_in=fan_out=o
    (ops/layers.py:52-53 initialise the bias with the same initializer)."""
    if len(shape) == 1:
        fan_in = fan_out = shape[0]
    elif len(shape) == 2:
        fan_in, fan_out = shape
    else:
        rf = 1
        for d in shape[:-2]:
            rf *= d
        fan_in, fan_out = rf * shape[-2], rf * shape[-1]
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1).mul_(lim).to(dtype)


# --------------------------------------------------------------------------- optimizers (util.py:150-183)
def adam_step(p, g, m, v, t, lr, beta1, beta2, eps=1e-8):
    """tf.train.AdamOptimizer (A.4), step t >= 1: lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    m,v EMA; p -= lr_t*m/(sqrt(v)+eps) ('epsilon-hat' form).  In place."""
    lr_t = lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    p.sub_(lr_t * m / (v.sqrt() + eps))


def rmsprop_step(p, g, ms, mom, lr, decay=0.9, momentum=0.01, eps=1e-10):
    """tf.train.RMSPropOptimizer, not centered (A.4); `ms` starts at 1.0.  In place."""
    ms.mul_(decay).addcmul_(g, g, value=1 - decay)
    mom.mul_(momentum).add_(lr * g / torch.sqrt(ms + eps))
    p.sub_(mom)


def centered_rmsprop_step(p, g, ms, mg, mom, lr, decay=0.9, momentum=0.01, eps=1e-10):
    """tf.train.RMSPropOptimizer(centered=True) (util.py:160-163 with --centered): mg = mean gradient;
    the denominator is sqrt(ms - mg^2 + eps).  In place."""
    ms.mul_(decay).addcmul_(g, g, value=1 - decay)
    mg.mul_(decay).add_(g, alpha=1 - decay)
    mom.mul_(momentum).add_(lr * g / torch.sqrt(ms - mg * mg + eps))
    p.sub_(mom)


def adagrad_step(p, g, accum, lr):
    """tf.train.AdagradOptimizer (util.py:166-167), accumulator init 0.1; tf.train.ProximalAdagradOptimizer
    (util.py:173-174) with its default l1 = l2 = 0 is the same update.  In place."""
    accum.addcmul_(g, g)
    p.sub_(lr * g / accum.sqrt())


def adadelta_step(p, g, accum, accum_update, lr, rho=0.95, eps=1e-8):
    """tf.train.AdadeltaOptimizer(lr) (util.py:164-165), rho 0.95, epsilon 1e-8 (ApplyAdadelta).  In place."""
    accum.mul_(rho).addcmul_(g, g, value=1 - rho)
    upd = torch.sqrt(accum_update + eps) / torch.sqrt(accum + eps) * g
    accum_update.mul_(rho).addcmul_(upd, upd, value=1 - rho)
    p.sub_(lr * upd)


def ftrl_step(p, g, accum, linear, lr):
    """tf.train.FtrlOptimizer(lr) (util.py:182-183): learning_rate_power -0.5, accumulator init 0.1,
    l1 = l2 = 0 (ApplyFtrl).  In place."""
    new_accum = accum + g * g
    linear.add_(g - (new_accum.sqrt() - accum.sqrt()) / lr * p)
    p.copy_(-linear / (new_accum.sqrt() / lr))
    accum.copy_(new_accum)


def sgd_step(p, g, lr):
    p.sub_(lr * g)


def momentum_step(p, g, acc, lr, momentum):
    """tf.train.MomentumOptimizer: acc = momentum*acc + g; p -= lr*acc."""
    acc.mul_(momentum).add_(g)
    p.sub_(lr * acc)


# --------------------------------------------------------------------------- losses
def sigmoid_ce(logits, labels):
    """tf.nn.sigmoid_cross_entropy_with_logits (A.5): max(x,0) - x*z + log(1+exp(-|x|))."""
    return torch.clamp(logits, min=0) - logits * labels + torch.log1p(torch.exp(-logits.abs()))


def rmse(x, y):
    """hem/ops/losses.py:10-15 — sqrt(mean((x-y)^2)); the one op the reference's own tests
    pin (hem/ops/test_losses.py:7-27)."""
    return torch.sqrt(torch.mean((x - y) ** 2))


def average_gradients(tower_grads):
    """util.py:118-147 — per-variable mean over towers; tower_grads: list (per tower) of
    lists of gradients in identical variable order."""
    return [torch.stack(gs, 0).mean(0) for gs in zip(*tower_grads)]
