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
        self.state = OrderedDict()      # non-trainable variables (batch norm moving averages): name -> StateVar
        self.state_buf = None
        self.collect = False            # True: layers register their outputs in `collections` (summary passes)
        self.collections = {}           # 'conv_layers' | 'dense_layers' -> [(name, Tensor)]  (ops/layers.py:60,105,146)

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
        self.collections = {}

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

    def get_state(self, name, shape, value, physical_shape=None):
        """A non-trainable fp32 variable (tf.contrib.layers.batch_norm's moving_mean / moving_variance)."""
        full = self.path(name)
        sv = self.state.get(full)
        if sv is None:
            if self.finalized:
                raise K.B200Error("state variable %s requested after the store was finalized" % full)
            sv = StateVar(full, tuple(shape if physical_shape is None else physical_shape), tuple(shape), float(value))
            self.state[full] = sv
        return sv

    def add_to_collection(self, key, name, tensor):
        """tf.add_to_collection('conv_layers' | 'dense_layers', h) of the layers (ops/layers.py:60,105,146): only while
        a summary pass is collecting, so that training passes do not pin their activations."""
        if self.collect:
            self.collections.setdefault(key, []).append((self.path(name), tensor))

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
        off = 0
        for sv in self.state.values():
            sv.offset = off
            off += (sv.numel + _ALIGN - 1) // _ALIGN * _ALIGN
        self.state_buf = torch.zeros(max(off, 1), dtype=torch.float32, device=device)
        for sv in self.state.values():
            sv.buf = self.state_buf[sv.offset:sv.offset + sv.numel]
            sv.logical(sv.buf).fill_(sv.value)       # padded channels stay 0: they are never read
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


class StateVar:
    def __init__(self, name, shape, logical_shape, value):
        self.name, self.shape, self.logical_shape, self.value = name, shape, logical_shape, value
        self.numel = int(math.prod(shape))
        self.buf = None

    def logical(self, flat):
        t = flat.reshape(self.shape)
        if self.logical_shape != self.shape:
            t = t[tuple(slice(0, d) for d in self.logical_shape)]
        return t


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


def ones_initializer():
    def init(shape, gen):
        return torch.ones(shape, dtype=torch.float32)

    return init


def zeros_initializer():
    def init(shape, gen):
        return torch.zeros(shape, dtype=torch.float32)

    return init


_ALIGN = 64     # elements; keeps every view 16-byte aligned in fp32 and bf16


# optimizer name -> (kernel kind, initial value of the `v` slot).  TF defaults: RMSProp's mean square starts
# at 1.0; Adagrad / ProximalAdagrad / FTRL accumulators at 0.1.
_OPTIMIZERS = {"adam": (K.OPT_ADAM, 0.0), "rmsprop": (K.OPT_RMSPROP, 1.0), "sgd": (K.OPT_SGD, 0.0),
               "momentum": (K.OPT_MOMENTUM, 0.0), "adagrad": (K.OPT_ADAGRAD, 0.1), "padagrad": (K.OPT_ADAGRAD, 0.1),
               "adadelta": (K.OPT_ADADELTA, 0.0), "ftrl": (K.OPT_FTRL, 0.1)}


class Group:
    """One optimizer instance's variables as flat buckets: p32 | g32 | m | v (fp32) and p16 (bf16)."""

    def __init__(self, name, params, cfg, host_values, device):
        self.name, self.params, self.cfg = name, list(params), dict(cfg)
        off = 0
        for p in self.params:
            p.offset = off
            off += (p.numel + _ALIGN - 1) // _ALIGN * _ALIGN
        self.size = off
        self.kind, v0 = _OPTIMIZERS[cfg["optimizer"]]
        if self.kind == K.OPT_RMSPROP and cfg.get("centered"):
            self.kind = K.OPT_CENTERED_RMSPROP
        self.p32 = torch.zeros(off, dtype=torch.float32, device=device)
        self.g32 = torch.zeros(off, dtype=torch.float32, device=device)
        self.m = torch.zeros(off, dtype=torch.float32, device=device)
        self.v = torch.full((off,), v0, dtype=torch.float32, device=device)
        # centered RMSProp keeps a third slot (the mean gradient)
        self.s3 = torch.zeros(off, dtype=torch.float32, device=device) if self.kind == K.OPT_CENTERED_RMSPROP else None
        self.p16 = torch.zeros(off, dtype=torch.bfloat16, device=device)
        self.step = torch.zeros(1, dtype=torch.int32, device=device)
        for p in self.params:
            sl = slice(p.offset, p.offset + p.numel)
            p.p32, p.g32, p.p16, p.group = self.p32[sl], self.g32[sl], self.p16[sl], self
            p.p32.copy_(host_values[p.name].reshape(-1))
            if p.need_t:
                p.p16_t = torch.zeros(p.numel, dtype=torch.bfloat16, device=device)
        self._t_table = None
        self.sync_compute_copies()

    # ---- optimizer slots by their TF names (checkpoints: `<var>/Adam`, `<var>/Adam_1`, `<var>/RMSProp`, ...)
    def slot_names(self):
        # TF-1.x names slot variables `<var>/<OptimizerName>[_k]` in creation order (rms, [mg,] momentum for RMSProp)
        return {K.OPT_ADAM: ("Adam", "Adam_1"), K.OPT_RMSPROP: ("RMSProp_1", "RMSProp"),
                K.OPT_CENTERED_RMSPROP: ("RMSProp_2", "RMSProp", "RMSProp_1"), K.OPT_SGD: (),
                K.OPT_MOMENTUM: ("Momentum",), K.OPT_ADAGRAD: (None, "Adagrad"),
                K.OPT_ADADELTA: ("Adadelta_1", "Adadelta"), K.OPT_FTRL: ("Ftrl_1", "Ftrl")}[self.kind]

    def slots(self):
        """[(tf slot name, flat buffer)] of the slots this optimizer uses (m, v, third)."""
        bufs = (self.m, self.v, self.s3)
        return [(n, b) for n, b in zip(self.slot_names(), bufs) if n is not None]

    def sync_compute_copies(self):
        self.p16.copy_(self.p32)
        if self.p32.is_cuda:
            self.refresh_transposed()

    def refresh_transposed(self):
        """Rebuild every per-tap transposed bf16 weight copy of the group (the K-major fprop operand) from the bf16
        compute copy: one launch over a device-resident table of (source, destination, T, A, B)."""
        need = [p for p in self.params if p.p16_t is not None]
        if not need or E.S.dry:
            return
        if self._t_table is None or self._t_table[2] != [id(p.p16_t) for p in need]:
            import ctypes as C
            ents = (K.TransposeEntry * len(need))()
            tiles = 0
            for e, p in zip(ents, need):
                a, b = p.shape[-2], p.shape[-1]
                t = p.numel // (a * b)
                e.src, e.dst, e.tile_begin, e.T, e.A, e.B = p.p16.data_ptr(), p.p16_t.data_ptr(), tiles, t, a, b
                tiles += t * ((a + 63) // 64) * ((b + 63) // 64)
            host = torch.frombuffer(bytearray(bytes(ents)), dtype=torch.uint8).clone()
            self._t_table = (host.to(self.p32.device), tiles, [id(p.p16_t) for p in need], len(need))
        tab, tiles, _, n = self._t_table
        E.launch("b200_transpose_batch", E._p(tab), n, tiles)

    def zero_grad(self):
        E.launch("b200_fill_f32", E._p(self.g32), self.size, 0.0)

    def apply_gradients(self, grad_scale=1.0, clip=0.0):
        """opt.apply_gradients (models/gan.py:80-81): one fused pass updates p / slots / the bf16 copy and resets
        the gradient bucket to zero for the next run; then the transposed weight copies the K-major fprop operand
        reads are re-laid out in one launch."""
        c = self.cfg
        kind = self.kind
        if kind == K.OPT_ADAM:
            b1, b2, eps = c["beta1"], c["beta2"], 1e-8
        elif kind in (K.OPT_RMSPROP, K.OPT_CENTERED_RMSPROP):
            b1, b2, eps = c["decay"], c["momentum"], 1e-10
        elif kind == K.OPT_MOMENTUM:
            b1, b2, eps = c["momentum"], 0.0, 0.0
        elif kind == K.OPT_ADADELTA:
            b1, b2, eps = 0.95, 0.0, 1e-8
        else:
            b1 = b2 = eps = 0.0
        E.launch("b200_optim_step", E._p(self.p32), E._p(self.m), E._p(self.v), E._p(self.s3), E._p(self.g32),
                 E._p(self.p16), self.size, kind, c["lr"], b1, b2, eps, grad_scale, clip, 1, E._p(self.step), n=2)
        self.refresh_transposed()


def optimizer_cfg(args):
    """init_optimizer(args) (util.py:150-183): the hyper-parameters of one optimizer instance."""
    name = args.optimizer
    if name == "pgd":
        raise K.B200Error("optimizer 'pgd': the reference's init_optimizer builds it without returning it "
                          "(util.py:171-172), so no reference run can use it")
    if name not in _OPTIMIZERS:
        raise K.B200Error("unknown optimizer '%s' (util.py:150-183 has: %s)" % (name, ", ".join(sorted(_OPTIMIZERS))))
    return {"optimizer": name, "lr": float(args.lr), "beta1": float(getattr(args, "beta1", 0.9)),
            "beta2": float(getattr(args, "beta2", 0.999)), "decay": float(getattr(args, "decay", 0.9)),
            "momentum": float(getattr(args, "momentum", 0.0)),
            "centered": bool(getattr(args, "centered", False)) and name == "rmsprop"}
