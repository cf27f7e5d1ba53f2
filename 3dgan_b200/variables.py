"""Variable store, TF-style scopes and optimizer buckets.

Mirrors how the reference's layers create variables (`tf.variable_scope('vars', reuse=reuse)` +
`tf.get_variable(name + '/weights')`, ops/layers.py:50-53) so checkpoints / parity fixtures can be
keyed by the reference's own variable names, and replaces `init_optimizer` + `apply_gradients`
(util.py:150-183, models/gan.py:80-81) by one fused update over a flat fp32 bucket per optimizer.
"""
import math
from collections import OrderedDict

import torch

from . import _capi as K
from . import engine as E


class VariableStore:
    def __init__(self, seed=0):
        self.params = OrderedDict()
        self.scope = []                 # name stack
        self.bn_counters = {}           # scope path -> next BatchNorm index (reset every pass)
        self.seed = seed
        self.finalized = False
        self.groups = []

    # ---------------------------------------------------------------- scopes
    def path(self, name=None):
        parts = [s for s in self.scope if s]
        if name:
            parts.append(name)
        return "/".join(parts)

    def begin_pass(self):
        """Call at the start of every graph-build / step pass: TF uniquifies BatchNorm scopes once at
        graph construction, so every replay of the model function must regenerate the same names."""
        self.bn_counters = {}
        self.scope = []

    def unique_bn_scope(self):
        base = self.path()
        k = self.bn_counters.get(base, 0)
        self.bn_counters[base] = k + 1
        return "BatchNorm" if k == 0 else "BatchNorm_%d" % k

    def get_variable(self, name, shape, init, physical_shape=None):
        """`shape` is the TF variable's shape; `physical_shape` (>= shape per dim) is how it is stored when
        a layer zero-pads its channels: the extra entries are initialised to zero and, receiving zero
        gradients, stay zero under every optimizer here."""
        full = self.path(name)
        p = self.params.get(full)
        phys = tuple(shape if physical_shape is None else physical_shape)
        if p is None:
            if self.finalized:
                raise K.B200Error("variable %s requested after the store was finalized" % full)
            p = E.Param(full, phys, logical_shape=shape)
            p.init = init
            self.params[full] = p
        elif tuple(shape) != p.logical_shape or phys != p.shape:
            raise ValueError("variable %s: shape %s != existing %s" % (full, tuple(shape), p.logical_shape))
        return p

    def collection(self, prefix):
        """tf.get_collection(TRAINABLE_VARIABLES, scope) (models/gan.py:65-66)."""
        return [p for n, p in self.params.items() if n.startswith(prefix)]

    # ---------------------------------------------------------------- materialise
    def finalize(self, groups, device):
        """groups: list of (name, [Param...], optimizer_cfg).  Allocates the flat buckets, initialises
        every variable (xavier-uniform for weights AND biases) and builds bf16 copies."""
        gen = torch.Generator().manual_seed(self.seed)
        host = OrderedDict()
        for name, p in self.params.items():           # creation order, like TF's initializer run
            host[name] = _pad_to(p.init(p.logical_shape, gen), p.shape)
        seen = set()
        for gname, params, cfg in groups:
            for p in params:
                assert id(p) not in seen, "variable %s is in two optimizer groups" % p.name
                seen.add(id(p))
            self.groups.append(Group(gname, params, cfg, host, device))
        missing = [n for n, p in self.params.items() if id(p) not in seen]
        if missing:
            raise K.B200Error("variables without an optimizer group: %s" % missing[:4])
        self.finalized = True

    def load(self, values):
        """Overwrite variables from {tf_name: tensor} (parity tests, checkpoints)."""
        for name, v in values.items():
            p = self.params[name]
            p.p32.copy_(_pad_to(v.reshape(p.logical_shape).to(torch.float32), p.shape).reshape(-1))
        for g in self.groups:
            g.sync_compute_copies()

    def state_dict(self):
        return OrderedDict((n, p.logical(p.p32.detach()).cpu().clone()) for n, p in self.params.items())


def _pad_to(t, shape):
    """zero-pad tensor t (logical shape) to `shape` (leading-block placement)."""
    if tuple(t.shape) == tuple(shape):
        return t
    out = torch.zeros(shape, dtype=t.dtype, device=t.device)
    out[tuple(slice(0, d) for d in t.shape)] = t
    return out


def xavier_initializer():
    """tf.contrib.layers.xavier_initializer() — returns init(shape, generator)."""

    def init(shape, gen):
        if len(shape) == 1:
            fan_in = fan_out = shape[0]
        elif len(shape) == 2:
            fan_in, fan_out = shape
        else:
            rf = int(math.prod(shape[:-2]))
            fan_in, fan_out = rf * shape[-2], rf * shape[-1]
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        return ((torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * lim).to(torch.float32)

    return init


def random_normal_initializer(mean=0.0, stddev=0.02):
    def init(shape, gen):
        return (torch.randn(shape, generator=gen, dtype=torch.float64) * stddev + mean).to(torch.float32)

    return init


def zeros_initializer():
    def init(shape, gen):
        return torch.zeros(shape, dtype=torch.float32)

    return init


_ALIGN = 64     # elements; keeps every view 16-byte aligned in fp32 and bf16


class Group:
    """One optimizer instance's variables as flat buckets: p32 | g32 | m | v (fp32) and p16 (bf16)."""

    def __init__(self, name, params, cfg, host_values, device):
        self.name, self.params, self.cfg = name, list(params), dict(cfg)
        off = 0
        for p in self.params:
            p.offset = off
            off += (p.numel + _ALIGN - 1) // _ALIGN * _ALIGN
        self.size = off
        self.p32 = torch.zeros(off, dtype=torch.float32, device=device)
        self.g32 = torch.zeros(off, dtype=torch.float32, device=device)
        self.m = torch.zeros(off, dtype=torch.float32, device=device)
        # RMSProp's mean-square slot starts at 1.0
        self.v = torch.full((off,), 1.0 if cfg["optimizer"] == "rmsprop" else 0.0, dtype=torch.float32,
                            device=device)
        self.p16 = torch.zeros(off, dtype=torch.bfloat16, device=device)
        self.step = torch.zeros(1, dtype=torch.int32, device=device)
        for p in self.params:
            sl = slice(p.offset, p.offset + p.numel)
            p.p32, p.g32, p.p16, p.group = self.p32[sl], self.g32[sl], self.p16[sl], self
            p.p32.copy_(host_values[p.name].reshape(-1))
            if p.need_t:
                p.p16_t = torch.zeros(p.numel, dtype=torch.bfloat16, device=device)
        self.sync_compute_copies()

    def sync_compute_copies(self):
        self.p16.copy_(self.p32)
        if self.p32.is_cuda:
            for p in self.params:
                p.refresh_transposed()

    def zero_grad(self):
        E.launch("b200_fill_f32", E._p(self.g32), self.size, 0.0)

    def apply_gradients(self, grad_scale=1.0, clip=0.0):
        """opt.apply_gradients (models/gan.py:80-81): fused update + bf16 copy, then re-layout the
        transposed weight copies the K-major fprop operand reads."""
        c = self.cfg
        kind = {"adam": K.OPT_ADAM, "rmsprop": K.OPT_RMSPROP, "sgd": K.OPT_SGD, "momentum": K.OPT_MOMENTUM}[
            c["optimizer"]]
        if kind == K.OPT_ADAM:
            b1, b2, eps = c["beta1"], c["beta2"], 1e-8
        elif kind == K.OPT_RMSPROP:
            b1, b2, eps = c["decay"], c["momentum"], 1e-10
        elif kind == K.OPT_MOMENTUM:
            b1, b2, eps = c["momentum"], 0.0, 0.0
        else:
            b1 = b2 = eps = 0.0
        E.launch("b200_optim_step", E._p(self.p32), E._p(self.m), E._p(self.v), E._p(self.g32), E._p(self.p16),
                 self.size, kind, c["lr"], b1, b2, eps, grad_scale, clip, E._p(self.step), n=2)
        for p in self.params:
            p.refresh_transposed()


def optimizer_cfg(args):
    """init_optimizer(args) (util.py:150-183): the hyper-parameters of one optimizer instance."""
    name = args.optimizer
    if name not in ("adam", "rmsprop", "sgd", "momentum"):
        raise K.B200Error("optimizer '%s' is not on the accelerated path (adam, rmsprop, sgd, momentum)" % name)
    if name == "rmsprop" and getattr(args, "centered", False):
        raise K.B200Error("centered RMSProp is not on the accelerated path")
    return {"optimizer": name, "lr": float(args.lr), "beta1": float(getattr(args, "beta1", 0.9)),
            "beta2": float(getattr(args, "beta2", 0.999)), "decay": float(getattr(args, "decay", 0.9)),
            "momentum": float(getattr(args, "momentum", 0.0))}
