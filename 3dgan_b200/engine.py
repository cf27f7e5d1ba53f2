"""Define-by-run tape over device buffers: the host side of the hot path.

The reference builds a symbolic TF graph from `ops/layers.py` calls and lets TF autodiff derive
the backward (models/gan.py:65-68, 224-231).  Here the same layer calls launch the CUDA kernels
immediately (through the C ABI in `_capi`) and record a tape node whose backward is written in
terms of the same ops, so second-order gradients (the IWGAN gradient penalty) come from running
the backward with recording on.  torch is used for buffers and streams only.

Mask convention: a tensor `t` may carry `t.mask = (src, kind, leak)` meaning
t = act(raw) with src = t, or t = raw * act'(src).  Whoever delivers a gradient to `t` multiplies
it by act'(src) (fused into the producing kernel's epilogue wherever possible), so every node's
backward receives the gradient w.r.t. its raw (pre-activation) output.
"""
import ctypes as C
import math
import os
import weakref

import torch

from . import _capi as K

BF16, F32 = torch.bfloat16, torch.float32


class _State:
    recording = False          # tape on/off
    accumulate = True          # backward sweeps add parameter gradients into the buckets
    dry = False                # build pass: shapes and variables only, no launches (meta tensors)
    active = frozenset()       # ids of Param objects whose gradients are being computed
    seq = 0
    stream = None              # ctypes stream handle for the current step
    device = None
    launches = 0               # kernels launched through the C ABI (bench.py reports it)
    profile = None             # list -> per-launch CUDA-event records of the conv-family launches
    decisions = None           # list -> (kind, tensors) of every discontinuous decision of the forward (tests)
    cyclic = []                # weak refs to tensors that sit on a reference cycle (self-mask, tape node)
    touched = {}               # id -> Param recorded on the tape since the last sweep that consumed them
    f32_out = False            # layer outputs are kept in fp32 (loss heads of the autoencoders)
    bn_updates = False         # this run executes batch norm's UPDATE_OPS (moving averages), models/gan.py:69-70


S = _State()


def _p(t):
    return None if (t is None or S.dry) else C.c_void_p(t.data_ptr())


def begin(device=None):
    """Bind the engine to torch's current device/stream for the calls that follow.  The previous step's tape is
    dropped: its tensors reference themselves (`t.mask = (t, ...)`) and their nodes, which would otherwise keep
    the step's activations alive until Python's cyclic collector runs."""
    S.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    S.stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    release_tape()


def release_tape():
    S.touched = {}
    live, S.cyclic = S.cyclic, []
    for r in live:
        t = r()
        if t is not None:
            t.mask = None
            t.node = None


def launch(name, *args, n=1, flops=0, tag=None):
    if S.dry:
        return 0
    S.launches += n
    if S.profile is not None and tag is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = K.call(name, *args, S.stream)
        e1.record()
        S.profile.append((name, flops, e0, e1, tag))
        return rc
    return K.call(name, *args, S.stream)


# relu/lrelu outputs also get a 1-bit/element sign map that the gradient epilogues read (B200GAN_SIGN_BITS=0: off)
SIGN_BITMAPS = os.environ.get("B200GAN_SIGN_BITS", "1") != "0"
SMALL_CHANNEL_GEMM = True      # False: image-side layers use the SIMT kernels (no workspace)
TAP_SPLIT = True               # False: never hand a split-K workspace to the tensor-core launches


def _workspace(g, op):
    """Scratch the call asks for (torch tensor as a byte buffer): im2col / col2im of the small-channel route, or the
    fp32 partial-sum image of a split-K launch (tensor-core layers with few output tiles)."""
    n = K.workspace_bytes(g, op)
    if n == 0 or not (SMALL_CHANNEL_GEMM if K.route(g, op) == 2 else TAP_SPLIT):
        return None, 0
    return empty((n,), torch.uint8), n


def _geom_key(g):
    return (g.N, g.H, g.W, g.Cin, g.Ho, g.Wo, g.Cout, g.k, g.stride, g.pad_t, g.pad_l)


def _conv_tag(op, g):
    """(tag, algorithmic FLOPs) of one conv-family launch: 2*N*Ho*Wo*k^2*Cin*Cout, logical channels
    (SURVEY 8d); tc: tensor-core route, simt: small-channel route."""
    opi = {"fprop": 0, "dgrad": 1, "wgrad": 2}[op]
    fam = "tc" if K.route(g, opi) == 1 else ("smallc-gemm" if SMALL_CHANNEL_GEMM and K.workspace_bytes(g, opi) else "simt")
    cin, cout = getattr(g, "logical", (g.Cin, g.Cout))       # zero-padded channels do no algorithmic work
    tag = "%s:%s N%d %dx%dx%d->%dx%dx%d k%ds%d" % (fam, op, g.N, g.H, g.W, cin, g.Ho, g.Wo, cout, g.k, g.stride)
    return tag, 2.0 * g.N * g.Ho * g.Wo * g.k * g.k * cin * cout


def empty(shape, dtype=BF16):
    if not S.dry and S.device is None:
        raise K.B200Error("engine.begin() has not bound a CUDA device / stream yet (Session.begin_step does it)")
    return torch.empty(shape, dtype=dtype, device="meta" if S.dry else S.device)


class f32_outputs:
    """Layers built inside this context store their output in fp32 instead of bf16.  The autoencoders use it
    for the final tanh / sigmoid layer, whose values feed the reconstruction loss directly: rounding a
    sigmoid output near 0 or 1 to bf16 destroys the Bernoulli loss gradient there (models/vae.py:76-79)."""

    def __enter__(self):
        self.prev, S.f32_out = S.f32_out, True

    def __exit__(self, *a):
        S.f32_out = self.prev


def _decision(kind, *tensors):
    """Verification hook (tests/parity.py): remember where the forward pass took a discontinuous decision —
    the sign masks of relu / lrelu outputs, the sign of an L1 residual — so that the oracle can be evaluated
    on the same linear piece (oracle.tf_ops.inject_decisions)."""
    if S.decisions is not None and not S.dry:
        S.decisions.append((kind,) + tensors)


class _All(frozenset):
    """Sentinel active-set: every variable is trainable (single-optimizer models)."""

    def __contains__(self, item):
        return True


ALL = _All()


class recording:
    def __init__(self, on=True, active=None):
        self.on, self.active = on, active

    def __enter__(self):
        self.prev = (S.recording, S.active)
        S.recording = self.on
        if self.active is not None:
            S.active = ALL if self.active == 'all' else frozenset(id(p) for p in self.active)
        return self

    def __exit__(self, *a):
        S.recording, S.active = self.prev


# ------------------------------------------------------------------------------------------ tensors
class Tensor:
    __slots__ = ("buf", "shape", "requires_grad", "_mask", "node", "grad_f32", "im2col", "bits", "logical_c",
                 "layout", "__weakref__")

    def __init__(self, buf, shape=None, requires_grad=False, mask=None):
        self.buf = buf
        self.shape = tuple(buf.shape if shape is None else shape)
        self.requires_grad = requires_grad
        self._mask = mask
        self.node = None
        self.grad_f32 = False      # leaf whose gradient is wanted in fp32 (the GP interpolates)
        self.im2col = None         # (geometry key, workspace) left by a small-channel fprop of this tensor
        self.bits = None           # int16 [rows, ceil(C/16)] sign bitmap written by the producing relu/lrelu epilogue
        self.logical_c = None      # channels that carry data when the last dim is zero-padded (ops/layers.py)
        self.layout = "NHWC"       # "NCHW": an input batch as the Gen-2 pipeline yields it (converted by the first op)

    @property
    def mask(self):
        return self._mask

    @mask.setter
    def mask(self, m):
        self._mask = m
        if m is not None and m[0] is self:
            S.cyclic.append(weakref.ref(self))

    @property
    def f32(self):
        return self.buf.dtype == F32

    @property
    def numel(self):
        return self.buf.numel()

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return reshape(self, shape)

    def torch(self):
        return self.buf.view(self.shape)


class Node:
    __slots__ = ("seq", "inputs", "outputs", "bw", "active")


def _record(inputs, outputs, bw, uses_active_param=False, params=()):
    if not S.recording:
        return
    if not (uses_active_param or any(t.requires_grad for t in inputs)):
        return
    n = Node()
    S.seq += 1
    n.seq, n.inputs, n.outputs, n.bw, n.active = S.seq, list(inputs), list(outputs), bw, S.active
    for p in params:
        # earliest tape position that uses this variable: once a reverse sweep has passed it, the variable's
        # gradient is final (Session.Exchange starts its all-reduce there, under the rest of the backward)
        if p is not None and p.active and id(p) not in S.touched:
            p.first_seq = n.seq
            S.touched[id(p)] = p
    for o in outputs:
        o.requires_grad = True
        o.node = n
        S.cyclic.append(weakref.ref(o))


class Param:
    """One trainable variable: fp32 master / gradient views into the group's flat buckets, a bf16
    compute copy, and (lazily) a per-tap transposed bf16 copy for the K-major fprop operand."""

    def __init__(self, name, shape, logical_shape=None):
        self.name, self.shape = name, tuple(shape)
        # physical shape may be zero-padded along the channel dims; the TF variable is the leading block
        self.logical_shape = tuple(shape if logical_shape is None else logical_shape)
        self.numel = int(math.prod(shape))
        self.p32 = self.g32 = self.p16 = self.p16_t = None
        self.group = None
        self.need_t = False
        self.first_seq = None

    @property
    def active(self):
        return id(self) in S.active

    @property
    def accum(self):
        """True when the running backward sweep should add this variable's gradient to its bucket."""
        if not S.accumulate or id(self) not in S.active:
            return False
        return S.accumulate is True or id(self) in S.accumulate

    def logical(self, flat):
        """The TF variable's view of one of this variable's flat buffers (p32 / g32 / ...)."""
        t = flat.reshape(self.shape)
        if self.logical_shape != self.shape:
            t = t[tuple(slice(0, d) for d in self.logical_shape)]
        return t

    def transposed(self):
        """bf16 [T, B, A] copy of a [T, A, B] weight (T = k*k taps)."""
        self.need_t = True
        if S.dry:
            return None
        if self.p16_t is None:
            self.p16_t = empty((self.numel,), BF16)
            self.refresh_transposed()
        return self.p16_t

    def refresh_transposed(self):
        """This variable alone (lazy creation, stand-alone test Params); groups refresh all of theirs in one launch."""
        if self.p16_t is None:
            return
        a, b = self.shape[-2], self.shape[-1]
        t = self.numel // (a * b)
        launch("b200_transpose_to_bf16", _p(self.p32), 1, _p(self.p16_t), t, a, b)
        if self.group is not None:
            self.group._t_table = None


def _act_code(act):
    return {None: K.ACT_NONE, "none": K.ACT_NONE, "relu": K.ACT_RELU, "lrelu": K.ACT_LRELU, "tanh": K.ACT_TANH,
            "sigmoid": K.ACT_SIGMOID}[act] if not isinstance(act, int) else act


def _epilogue(bias=None, act=K.ACT_NONE, leak=0.0, mask=None, out_f32=False, accumulate=False, bits_ok=False,
              out_c=None):
    e = K.Epilogue()
    e.bias = None if (bias is None or S.dry) else bias.data_ptr()
    e.act, e.leak = act, leak
    if mask is not None:
        e.mask_src, e.mask_kind = (None if S.dry else mask[0].buf.data_ptr()), mask[1]
        # the bitmap is laid out per row of C channels: usable only when the mask tensor has this output's
        # row width (a mask seen through a reshape keeps its values but not its bitmap layout)
        if (bits_ok and mask[0].bits is not None and mask[1] in (K.ACT_RELU, K.ACT_LRELU)
                and mask[0].shape[-1] == out_c):
            e.mask_src, e.mask_bits, e.bits_pitch = None, mask[0].bits.data_ptr(), mask[0].bits.shape[1]
        if act == K.ACT_NONE:
            e.leak = mask[2]
        elif mask[1] == K.ACT_LRELU and act == K.ACT_LRELU and mask[2] != leak:
            raise K.B200Error("epilogue cannot fuse two different lrelu leaks")
    e.out_f32, e.accumulate = int(out_f32), int(accumulate)
    return e


def same_pad(size, k, s):
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return out, total // 2


def conv_geom(N, H, W, Cin, Cout, k, stride, padding='SAME'):
    """TF geometry of conv [N,H,W,Cin] -> [N,Ho,Wo,Cout]: SAME (Ho = ceil(H/s), the smaller half of the
    padding first) or VALID (no padding, Ho = (H - k) // s + 1; hem/ops/layers.py:118 `padding=padding`)."""
    g = K.ConvGeom()
    g.N, g.H, g.W, g.Cin, g.Cout, g.k, g.stride = N, H, W, Cin, Cout, k, stride
    if padding == 'SAME':
        g.Ho, g.pad_t = same_pad(H, k, stride)
        g.Wo, g.pad_l = same_pad(W, k, stride)
    elif padding == 'VALID':
        if H < k or W < k:
            raise K.B200Error("VALID convolution: the %dx%d filter does not fit a %dx%d input" % (k, k, H, W))
        g.Ho, g.Wo, g.pad_t, g.pad_l = (H - k) // stride + 1, (W - k) // stride + 1, 0, 0
    else:
        raise K.B200Error("padding must be 'SAME' or 'VALID'")
    return g


# ------------------------------------------------------------------------------------------ conv family
def conv_like(direction, x, W, geom, bias=None, act=K.ACT_NONE, leak=0.0, out_mask=None, out_f32=False):
    """direction 'fprop': x [N,H,W,Cin] -> [N,Ho,Wo,Cout];  'dgrad': x [N,Ho,Wo,Cout] -> [N,H,W,Cin]
    (the adjoint: TF's conv2d_transpose / input gradient).  W is a Param laid out [k,k,Cin,Cout].
    out = act(op(x,W) + bias) * act'(out_mask.src)."""
    g = geom
    if x.f32:
        raise K.B200Error("conv_like needs a bf16 input")
    if direction == "fprop":
        out_shape = (g.N, g.Ho, g.Wo, g.Cout)
    else:
        out_shape = (g.N, g.H, g.W, g.Cin)
    out_f32 = out_f32 or S.f32_out
    out = Tensor(empty(out_shape, F32 if out_f32 else BF16))
    opi = 0 if direction == "fprop" else 1
    ws, wsb = _workspace(g, opi)
    bits_ok = SIGN_BITMAPS and not S.dry and K.epilogue_bits(g, opi, ws is not None)
    e = _epilogue(None if bias is None else bias.p32, act, leak, out_mask, out_f32, bits_ok=bits_ok,
                  out_c=out_shape[-1])
    if bits_ok and act in (K.ACT_RELU, K.ACT_LRELU) and not out_f32:
        c = out_shape[-1]
        out.bits = torch.empty((out.numel // c, (c + 15) // 16), dtype=torch.int16, device=out.buf.device)
        e.bits_out, e.bits_pitch = out.bits.data_ptr(), out.bits.shape[1]
    if direction == "fprop":
        wt = W.transposed() if K.route(g, 0) == 1 else None
        tag, fl = _conv_tag("fprop", g) if S.profile is not None else (None, 0)
        launch("b200_conv2d_fprop", _p(x.buf), _p(W.p16), _p(wt), _p(out.buf), C.byref(g), C.byref(e), _p(ws), wsb,
               flops=fl, tag=tag)
        if ws is not None and wt is None:
            x.im2col = (_geom_key(g), ws)       # the filter gradient of this layer reads the same im2col
    else:
        tag, fl = _conv_tag("dgrad", g) if S.profile is not None else (None, 0)
        launch("b200_conv2d_dgrad", _p(x.buf), _p(W.p16), _p(out.buf), C.byref(g), C.byref(e), _p(ws), wsb, flops=fl,
               tag=tag)
    if act != K.ACT_NONE:
        if out_mask is not None:
            raise K.B200Error("conv_like: activation and out_mask are mutually exclusive")
        out.mask = (out, act, leak)
        if act in (K.ACT_RELU, K.ACT_LRELU):
            _decision("act", out)
    else:
        out.mask = out_mask

    def bw(gouts):
        (go,) = gouts
        go = _as_bf16(go)
        gx = None
        if x.requires_grad:
            gx = conv_like("dgrad" if direction == "fprop" else "fprop", go, W, g, out_mask=x.mask,
                           out_f32=x.grad_f32)
        if W.accum:
            tag, fl = _conv_tag("wgrad", g) if S.profile is not None else (None, 0)
            a_, b_ = (x, go) if direction == "fprop" else (go, x)
            cached = a_.im2col if (a_.im2col is not None and a_.im2col[0] == _geom_key(g)) else None
            if cached is not None:
                ws, wsb, ready = cached[1], cached[1].numel(), 1
            else:
                (ws, wsb), ready = _workspace(g, 2), 0
            # image-side layers: the gathered rows carry a ones column, so the same pass yields the bias gradient
            fold = (bias is not None and bias.accum and direction == "fprop" and ws is not None
                    and K.wgrad_folds_bias(g, True))
            if fold:
                launch("b200_conv2d_wgrad_bias", _p(a_.buf), _p(b_.buf), _p(W.g32), _p(bias.g32), C.byref(g), 1.0,
                       _p(ws), wsb, ready, flops=fl, tag=tag)
                return [gx]
            launch("b200_conv2d_wgrad", _p(a_.buf), _p(b_.buf), _p(W.g32), C.byref(g), 1.0, _p(ws), wsb, ready,
                   flops=fl, tag=tag)
        if bias is not None and bias.accum:
            c = out_shape[-1]
            launch("b200_colsum", _p(go.buf), None, _p(bias.g32), go.numel // c, c, 1.0)
        return [gx]

    _record([x], [out], bw, W.active or (bias is not None and bias.active), params=(W, bias))
    return out


def _as_bf16(t):
    if not t.f32:
        return t
    o = Tensor(empty(t.shape, BF16))
    launch("b200_affine_act", _p(t.buf), 1, _p(o.buf), 0, t.numel, 1.0, 0.0, 0, 0.0)
    return o


# ------------------------------------------------------------------------------------------ dense, one unit
def dense_n1(x, W, bias, act=K.ACT_NONE, leak=0.0):
    """x [M,K] bf16, W Param [K,1] -> fp32 [M]  (critic fc2, models/gan.py:285)."""
    M, Kd = x.shape
    out = Tensor(empty((M,), F32))
    launch("b200_gemv_rows", _p(x.buf), _p(W.p16), _p(bias.p32), _p(out.buf), M, Kd, act, leak)
    if act != K.ACT_NONE:
        out.mask = (out, act, leak)

    def bw(gouts):
        (go,) = gouts           # fp32 [M]; if act != none the deliverer already applied act'
        gx = outer_mask(go, W, x) if x.requires_grad else None
        if W.accum:
            launch("b200_colsum", _p(x.buf), _p(go.buf), _p(W.g32), M, Kd, 1.0)
        if bias.accum:
            launch("b200_reduce_sum", _p(go.buf), 1, M, _p(bias.g32), 1.0, 0)
        return [gx]

    _record([x], [out], bw, W.active or bias.active, params=(W, bias))
    return out


def outer_mask(g, W, like):
    """out[m,k] = g[m] * W[k] * act'(like.mask) — gradient of dense_n1 w.r.t. its input."""
    M, Kd = like.shape
    out = Tensor(empty((M, Kd), BF16), mask=like.mask)
    m = like.mask
    launch("b200_outer_mask", _p(g.buf), _p(W.p16), None if m is None else _p(m[0].buf), _p(out.buf), M, Kd,
           0 if m is None else m[1], 0.0 if m is None else m[2])

    def bw(gouts):
        (c,) = gouts            # bf16 [M,K], already multiplied by act'(mask)
        if W.accum:
            launch("b200_colsum", _p(c.buf), _p(g.buf), _p(W.g32), M, Kd, 1.0)
        gg = None
        if g.requires_grad:
            gg = Tensor(empty((M,), F32))
            launch("b200_gemv_rows", _p(c.buf), _p(W.p16), None, _p(gg.buf), M, Kd, 0, 0.0)
        return [gg]

    _record([g], [out], bw, W.active, params=(W,))
    return out


# ------------------------------------------------------------------------------------------ batch norm
def batch_norm_act(z, beta, act=K.ACT_NONE, leak=0.0, eps=1e-3, moving=None, decay=0.999, unbiased=False):
    """tf.contrib.layers.batch_norm(h) with defaults, then the layer activation
    (ops/layers.py:58-59,103-104,144-145): beta only, biased batch statistics, eps 1e-3.
    moving = (moving_mean, moving_variance) fp32 buffers, updated when the run executes UPDATE_OPS."""
    Cc = z.shape[-1]
    R = z.numel // Cc
    stats = empty((2 * Cc,), F32)
    launch("b200_fill_f32", _p(stats), 2 * Cc, 0.0)
    launch("b200_bn_sums", _p(z.buf), _p(stats), R, Cc)
    if moving is not None and S.bn_updates:
        launch("b200_bn_update_moving", _p(stats), R, Cc, _p(moving[0]), _p(moving[1]), decay, int(unbiased))
    out = Tensor(empty(z.shape, BF16))
    launch("b200_bn_apply", _p(z.buf), _p(stats), _p(beta.p32), _p(out.buf), R, Cc, eps, act, leak)
    if act != K.ACT_NONE:
        out.mask = (out, act, leak)
        if act in (K.ACT_RELU, K.ACT_LRELU):
            _decision("act", out)

    def bw(gouts):
        (go,) = gouts
        go = _as_bf16(go)
        bsum = empty((2 * Cc,), F32)
        launch("b200_fill_f32", _p(bsum), 2 * Cc, 0.0)
        dz = Tensor(empty(z.shape, BF16))
        launch("b200_bn_bwd", _p(go.buf), _p(z.buf), _p(stats), _p(bsum), _p(dz.buf), R, Cc, eps, n=2)
        if beta.accum:
            launch("b200_axpby", _p(bsum), 1, 1.0, None, _p(beta.g32), 1, 1.0, _p(beta.g32), 1, Cc)
        if z.mask is not None:
            dz = maskmul(dz, z.mask)
        return [dz]

    _record([z], [out], bw, beta.active, params=(beta,))
    return out


# ------------------------------------------------------------------------------------------ elementwise
def maskmul(g, mask):
    out = Tensor(empty(g.shape, BF16), mask=mask)
    g = _as_bf16(g)
    launch("b200_maskmul", _p(g.buf), _p(mask[0].buf), _p(out.buf), g.numel, mask[1], mask[2])

    def bw(gouts):
        return [gouts[0]]       # deliverer already applied the mask (out.mask == mask)

    _record([g], [out], bw)
    return out


def activation(x, act, leak=0.0):
    """Stand-alone activation (used when a layer's activation cannot be fused)."""
    out = Tensor(empty(x.shape, BF16))
    launch("b200_affine_act", _p(x.buf), int(x.f32), _p(out.buf), 0, x.numel, 1.0, 0.0, act, leak)
    out.mask = (out, act, leak)
    out.logical_c = x.logical_c
    if act in (K.ACT_RELU, K.ACT_LRELU):
        _decision("act", out)

    def bw(gouts):
        go = gouts[0]
        return [maskmul(go, x.mask) if x.mask is not None else go]

    _record([x], [out], bw)
    return out


def affine(x, mul, add, out_f32=False):
    """out = x*mul + add (e.g. the [0,1] -> [-1,1] rescale, models/gan.py:50).  A uint8 input is an image batch as
    decoded (data.py:14-22): its /255 normalisation is folded into this pass."""
    if x.layout == "NCHW":
        # an input batch in the reference's Gen-2 layout (hem/ops/layers.py:117-119): transposed to the kernels' NHWC in
        # the same pass that rescales (and, for image bytes, normalises) it
        n, c, h, w = x.shape
        out = Tensor(empty((n, h, w, c), BF16))
        u8 = x.buf.dtype == torch.uint8
        launch("b200_layout_convert", _p(x.buf), 2 if u8 else int(x.f32), _p(out.buf), 0, n, c, h * w,
               mul / 255.0 if u8 else mul, add)
        return out
    out = Tensor(empty(x.shape, F32 if out_f32 else BF16))
    if x.buf.dtype == torch.uint8:
        launch("b200_affine_act", _p(x.buf), 2, _p(out.buf), int(out_f32), x.numel, mul / 255.0, add, 0, 0.0)
        return out                                   # an input batch: no gradient flows back to it
    launch("b200_affine_act", _p(x.buf), int(x.f32), _p(out.buf), int(out_f32), x.numel, mul, add, 0, 0.0)

    def bw(gouts):
        go = gouts[0]
        o = Tensor(empty(x.shape, BF16))
        launch("b200_affine_act", _p(go.buf), int(go.f32), _p(o.buf), 0, go.numel, mul, 0.0, 0, 0.0)
        return [maskmul(o, x.mask) if x.mask is not None else o]

    _record([x], [out], bw)
    return out


def reshape(x, shape):
    shape = tuple(shape)
    last = shape[-1]
    if x.logical_c is not None and last != x.shape[-1]:
        # the zero channel padding must not be mixed into the data when the last dim is re-cut
        # (models/gan.py:284: the critic's c3 output -> rows of 4*4*4L, any latent_size)
        x = unpad_channels(x)
    if -1 in shape:
        known = -int(math.prod(shape))
        if known <= 0 or x.numel % known:
            raise K.B200Error("reshape: cannot infer -1 in %s from %d elements" % (shape, x.numel))
        shape = tuple(x.numel // known if d == -1 else d for d in shape)
    if int(math.prod(shape)) != x.numel:
        raise K.B200Error("reshape: %s -> %s changes the element count" % (x.shape, shape))
    out = Tensor(x.buf, shape, mask=x.mask)
    if last == x.shape[-1]:
        out.logical_c = x.logical_c

    def bw(gouts):
        return [reshape(gouts[0], x.shape)]     # through the op so second-order tapes stay connected

    _record([x], [out], bw)
    return out


def _slice(src, src_ld, src_off, dst, dst_ld, dst_off, rows, cols, mask=None):
    launch("b200_slice_cols", _p(src), src_ld, src_off, _p(dst), dst_ld, dst_off, rows, cols,
           None if mask is None else _p(mask[0].buf), 0 if mask is None else mask[1], 0.0 if mask is None else mask[2])


def unpad_channels(x):
    """[..., C_physical] -> [..., C_logical]: drop the zero channels a layer added (ops/layers.py)."""
    cl = x.logical_c
    out = take_channels(x, cl)
    if x.mask is not None and x.mask[0] is x:
        out.mask = (out, x.mask[1], x.mask[2])          # same activation values, own storage
    return out


def take_channels(x, c, mask=None):
    """out = x[..., :c] * act'(mask): leading channel slice (un-padding), with the activation gradient of the
    tensor the result is a gradient FOR fused in.  Its backward zero-pads again (both directions are recorded
    ops, so the gradient-penalty's second-order sweep passes through them)."""
    cp = x.shape[-1]
    rows = x.numel // cp
    xb = _as_bf16(x)
    out = Tensor(empty(x.shape[:-1] + (c,), BF16), mask=mask)
    _slice(xb.buf, cp, 0, out.buf, c, 0, rows, c, mask)

    def bw(gouts):
        gx = pad_channels(gouts[0], cp)
        return [maskmul(gx, x.mask) if (x.mask is not None and x.mask[0] is not x) else gx]

    _record([x], [out], bw)
    return out


def pad_channels(x, cp):
    """[..., C] -> [..., cp] with zero channels appended (the inverse of unpad_channels): gives a dense layer whose
    input width is not a multiple of 8 (e.g. latent_size 50) the 16-byte rows TMA needs, and is the backward of
    take_channels."""
    c = x.shape[-1]
    rows = x.numel // c
    xb = _as_bf16(x)
    out = Tensor(empty(x.shape[:-1] + (cp,), BF16))
    launch("b200_fill_f32", _p(out.buf), out.numel // 2, 0.0)        # bf16 zeros, two per fp32 word
    _slice(xb.buf, c, 0, out.buf, cp, 0, rows, c)
    out.logical_c = c

    def bw(gouts):
        return [take_channels(gouts[0], c, x.mask)]

    _record([x], [out], bw)
    return out


def to_nchw(x, mul=1.0, add=0.0):
    """NHWC activation -> fp32 NCHW copy (samples / summaries handed back in the reference's Gen-2 layout)."""
    n, h, w, c = x.shape
    out = Tensor(empty((n, c, h, w), F32))
    out.layout = "NCHW"
    launch("b200_layout_convert", _p(x.buf), int(x.f32), _p(out.buf), 1, n, c, h * w, mul, add)
    return out


def add(a, b):
    """a + b (the shortcut of hem.residual, hem/ops/layers.py:304)."""
    out = Tensor(empty(a.shape, BF16))
    launch("b200_axpby", _p(a.buf), int(a.f32), 1.0, None, _p(b.buf), int(b.f32), 1.0, _p(out.buf), 0, a.numel)
    out.logical_c = a.logical_c

    def bw(gouts):
        go = gouts[0]
        return [maskmul(go, t.mask) if (t.requires_grad and t.mask is not None) else go for t in (a, b)]

    _record([a, b], [out], bw)
    return out


def dropout(x, keep_prob, u):
    """tf.nn.dropout(h, keep_prob) (hem/ops/layers.py:64,132,208; the reference passes its `dropout` value as
    keep_prob): out = x * [u >= 1 - keep] / keep, u ~ U[0,1) fp32 drawn by the caller (Session.random_uniform)."""
    xb = _as_bf16(x)
    out = Tensor(empty(x.shape, BF16))
    out.logical_c = x.logical_c
    launch("b200_dropout", _p(xb.buf), _p(u.buf), _p(out.buf), x.numel, float(keep_prob))

    def bw(gouts):
        go = _as_bf16(gouts[0])
        gx = Tensor(empty(x.shape, BF16))
        launch("b200_dropout", _p(go.buf), _p(u.buf), _p(gx.buf), x.numel, float(keep_prob))
        return [maskmul(gx, x.mask) if x.mask is not None else gx]

    _record([x], [out], bw)
    return out


def instance_norm(x, scale, shift, eps=1e-3):
    """hem.instance_norm (hem/ops/images.py:73-89): per sample and channel, moments over H x W; scale * xhat + shift."""
    n, h, w, c = x.shape
    xb = _as_bf16(x)
    stats = empty((n, 2 * c), F32)
    out = Tensor(empty(x.shape, BF16))
    out.logical_c = x.logical_c
    launch("b200_instnorm_fwd", _p(xb.buf), _p(scale.p32), _p(shift.p32), _p(out.buf), _p(stats), n, h * w, c, eps)

    def bw(gouts):
        go = _as_bf16(gouts[0])
        dx = Tensor(empty(x.shape, BF16))
        both = scale.accum and shift.accum
        ds = scale.g32 if both else empty((c,), F32)
        dh = shift.g32 if both else empty((c,), F32)
        if not both:
            launch("b200_fill_f32", _p(ds), c, 0.0)
            launch("b200_fill_f32", _p(dh), c, 0.0)
        launch("b200_instnorm_bwd", _p(go.buf), _p(xb.buf), _p(stats), _p(scale.p32), _p(dx.buf), _p(ds), _p(dh),
               n, h * w, c)
        return [maskmul(dx, x.mask) if x.mask is not None else dx]

    _record([x], [out], bw, scale.active or shift.active, params=(scale, shift))
    return out


def add_grads(a, b):
    out = Tensor(empty(a.shape, F32 if (a.f32 and b.f32) else BF16))
    launch("b200_axpby", _p(a.buf), int(a.f32), 1.0, None, _p(b.buf), int(b.f32), 1.0, _p(out.buf), int(out.f32),
           a.numel)
    return out


def concat_channels(xs):
    """NHWC channel concat (tf.concat(axis=1) of the NCHW reference: pix2pix skip connections and the PatchGAN's
    rgb+depth input, hem/models/pix2pix.py:210-222,250): every piece is copied into its column slice of the
    output; the backward reads the slice back and applies the piece's activation gradient in the same pass."""
    for t in xs:
        if t.logical_c is not None or t.f32:
            raise K.B200Error("concat_channels: pieces must be unpadded bf16 tensors")
    sizes = [t.shape[-1] for t in xs]
    ctot = sum(sizes)
    rows = xs[0].numel // sizes[0]
    out = Tensor(empty(xs[0].shape[:-1] + (ctot,), BF16))
    o = 0
    for t, c in zip(xs, sizes):
        if t.numel // c != rows:
            raise K.B200Error("concat_channels: pieces disagree on the leading dims")
        _slice(t.buf, c, 0, out.buf, ctot, o, rows, c)
        o += c

    def bw(gouts):
        go = _as_bf16(gouts[0])
        outs, o = [], 0
        for t, c in zip(xs, sizes):
            if t.requires_grad:
                piece = Tensor(empty(t.shape, BF16), mask=t.mask)
                _slice(go.buf, ctot, o, piece.buf, c, 0, rows, c, t.mask)
                outs.append(piece)
            else:
                outs.append(None)
            o += c
        return outs

    _record(xs, [out], bw)
    return out


def interpolate(x, g, alpha):
    """x_hat = x + alpha*(g - x), alpha [B] fp32 (models/gan.py:224-226).  A leaf for the tape:
    in the critic step G is a constant and in the generator step d_loss is only evaluated."""
    B = x.shape[0]
    out = Tensor(empty(x.shape, BF16))
    launch("b200_interp", _p(x.buf), _p(g.buf), _p(alpha.buf), _p(out.buf), B, x.numel // B)
    return out


def sumsq(x):
    """sum(x^2) -> fp32 [1] (tf.reduce_sum(tf.square(gradients)), models/gan.py:229)."""
    out = Tensor(empty((1,), F32))
    launch("b200_fill_f32", _p(out.buf), 1, 0.0)
    launch("b200_reduce_sum", _p(x.buf), int(x.f32), x.numel, _p(out.buf), 1.0, 1)

    def bw(gouts):
        go = gouts[0]           # fp32 [1] device scalar
        o = Tensor(empty(x.shape, BF16))
        launch("b200_axpby", _p(x.buf), int(x.f32), 2.0, _p(go.buf), None, 0, 0.0, _p(o.buf), 0, x.numel)
        return [maskmul(o, x.mask) if x.mask is not None else o]

    _record([x], [out], bw)
    return out


def wgan_losses(d_real, d_fake, ss=None, lam=10.0):
    """g_loss = -mean(d_fake); d_loss = mean(d_fake) - mean(d_real) [+ lam*(sqrt(ss)-1)^2]
    (models/gan.py:194-205).  Returns (g_loss, d_loss) fp32 [1] tensors sharing one node."""
    B = d_real.shape[0]
    sums = empty((3,), F32)
    launch("b200_fill_f32", _p(sums), 3, 0.0)
    launch("b200_reduce_sum", _p(d_real.buf), 1, B, _p(sums[0:1]), 1.0, 0)
    launch("b200_reduce_sum", _p(d_fake.buf), 1, d_fake.shape[0], _p(sums[1:2]), 1.0, 0)
    if ss is not None:
        launch("b200_axpby", _p(ss.buf), 1, 1.0, None, None, 0, 0.0, _p(sums[2:3]), 1, 1)
    out4 = empty((4,), F32)
    launch("b200_wgan_loss", _p(sums), B, int(ss is not None), lam, _p(out4))
    g_loss, d_loss = Tensor(out4[0:1]), Tensor(out4[1:2])
    inputs = [d_real, d_fake] + ([ss] if ss is not None else [])

    def bw(gouts):
        gg, gd = gouts
        cr = cf = None
        if gd is not None:       # seed 1.0 on d_loss
            cr = Tensor(empty((B,), F32)); cf = Tensor(empty((B,), F32))
            launch("b200_fill_f32", _p(cr.buf), B, -1.0 / B)
            launch("b200_fill_f32", _p(cf.buf), B, 1.0 / B)
        if gg is not None:       # seed 1.0 on g_loss
            t = Tensor(empty((B,), F32))
            launch("b200_fill_f32", _p(t.buf), B, -1.0 / B)
            cf = t if cf is None else add_grads(cf, t)
        res = [cr if d_real.requires_grad else None, cf if d_fake.requires_grad else None]
        if ss is not None:
            res.append(Tensor(out4[3:4]) if gd is not None else None)
        return res

    _record(inputs, [g_loss, d_loss], bw)
    return g_loss, d_loss


def reparameterize(mu, sd, eps):
    """z = mu + sd*eps (models/vae.py:127-128); eps is noise (no gradient)."""
    out = Tensor(empty(mu.shape, BF16))
    launch("b200_mul_add", _p(mu.buf), _p(sd.buf), _p(eps.buf), _p(out.buf), mu.numel)

    def bw(gouts):
        go = _as_bf16(gouts[0])
        gsd = Tensor(empty(sd.shape, BF16))
        launch("b200_mul_add", None, _p(go.buf), _p(eps.buf), _p(gsd.buf), sd.numel)
        gmu = maskmul(go, mu.mask) if mu.mask is not None else go
        if sd.mask is not None:
            gsd = maskmul(gsd, sd.mask)
        return [gmu, gsd]

    _record([mu, sd], [out], bw)
    return out


def add_scalars(a, b):
    """a + b for fp32 [1] loss pieces."""
    out = Tensor(empty((1,), F32))
    launch("b200_axpby", _p(a.buf), 1, 1.0, None, _p(b.buf), 1, 1.0, _p(out.buf), 1, 1)

    def bw(gouts):
        return [gouts[0], gouts[0]]

    _record([a, b], [out], bw)
    return out


def scale_scalar(a, s):
    """s * a for an fp32 [1] loss piece."""
    out = Tensor(empty((1,), F32))
    launch("b200_axpby", _p(a.buf), 1, float(s), None, None, 0, 0.0, _p(out.buf), 1, 1)

    def bw(gouts):
        go = gouts[0]
        if go is True:
            return [("scaled", float(s))]
        raise K.B200Error("scale_scalar: only a seed gradient is supported")

    _record([a], [out], bw)
    return out


def eltloss(a, b, kind, label=0.0, scale=1.0):
    """sum_i l(a_i, b_i) * scale -> fp32 [1]; backward fuses dl/da (kinds: see simt_kernels.cu)."""
    out = Tensor(empty((1,), F32))
    launch("b200_fill_f32", _p(out.buf), 1, 0.0)
    launch("b200_eltloss", _p(a.buf), int(a.f32), None if b is None else _p(b.buf), a.numel, kind, label, scale, 0.0,
           _p(out.buf), None, 0, 0, 0.0)
    if kind == 0:
        _decision("l1", a, b)

    def bw(gouts):
        seed = gouts[0]
        mult = seed[1] if isinstance(seed, tuple) else 1.0      # True = seed 1.0; ("scaled", s) = seed s
        g = Tensor(empty(a.shape, F32 if a.f32 else BF16))
        # a = act(raw) produced by the layer itself: dl/d(raw) = dl/da * act'(a) in one pass, in fp32 (the two
        # factors of the Bernoulli loss behind a sigmoid are ~1/(1-a) and ~(1-a): they must not be rounded apart)
        own = a.mask is not None and a.mask[0] is a
        launch("b200_eltloss", _p(a.buf), int(a.f32), None if b is None else _p(b.buf), a.numel, kind, label, 0.0,
               scale * mult, None, _p(g.buf), int(a.f32), a.mask[1] if own else 0, a.mask[2] if own else 0.0)
        if a.mask is not None and not own:
            g = maskmul_any(g, a.mask)
        return [g]

    _record([a], [out], bw)
    return out


def maskmul_any(g, mask):
    """maskmul that also accepts fp32 gradients/sources (tiny tensors only: goes through bf16)."""
    src = mask[0]
    if src.f32:
        src16 = _as_bf16(src)
        mask = (src16, mask[1], mask[2])
    o = maskmul(g, mask)
    if g.f32:
        f = Tensor(empty(g.shape, F32))
        launch("b200_affine_act", _p(o.buf), 0, _p(f.buf), 1, o.numel, 1.0, 0.0, 0, 0.0)
        return f
    return o


def random_fill(shape, normal, seed, counter, stream_id, f32=True):
    """tf.random_normal / tf.random_uniform stand-in (Philox4x32-10, device-side draw counter)."""
    out = Tensor(empty(shape, F32 if f32 else BF16))
    launch("b200_philox", _p(out.buf), int(f32), out.numel, seed, _p(counter), stream_id, int(normal), n=2)
    return out


# ------------------------------------------------------------------------------------------ backward pass
def backward(seeds, wrt=(), create_graph=False, accumulate=True, on_ready=None):
    """Reverse sweep.  seeds: list of (tensor, grad Tensor | None); None = seed 1.0 on a scalar loss.
    With accumulate=True (or a list of Params) parameter gradients are added into those Params' g32 views;
    tf.gradients(ys, xs)-style calls (models/gan.py:228) pass accumulate=False and read `wrt`.
    on_ready(param): called as soon as the sweep has passed the earliest tape node that uses `param`, i.e. when
    its gradient is final (every variable recorded since the last consuming sweep is reported, the rest at the end)."""
    grads = {}
    nodes = {}
    if accumulate not in (True, False):
        accumulate = frozenset(id(p) for p in accumulate)
    prev_acc, S.accumulate = S.accumulate, accumulate
    prev_active = S.active
    waiting = []
    if on_ready is not None and accumulate is not False:
        waiting = sorted((p for p in S.touched.values() if accumulate is True or id(p) in accumulate),
                         key=lambda p: -p.first_seq)

    def push(t, g):
        if t.node is None and not any(t is w for w in wrt):
            return
        key = id(t)
        if key in grads:
            is_seed = lambda v: v is True or isinstance(v, tuple)
            grads[key] = (t, add_grads(grads[key][1], g) if not is_seed(g) and not is_seed(grads[key][1]) else g)
        else:
            grads[key] = (t, g)
        if t.node is not None:
            nodes[t.node.seq] = t.node

    try:
        for t, g in seeds:
            push(t, True if g is None else g)
        with recording(create_graph):
            while nodes:
                seq = max(nodes)
                while waiting and waiting[0].first_seq > seq:
                    S.touched.pop(id(waiting[0]), None)
                    on_ready(waiting.pop(0))
                node = nodes.pop(seq)
                gouts = []
                for o in node.outputs:
                    ent = grads.pop(id(o), None) if not any(o is w for w in wrt) else grads.get(id(o))
                    gouts.append(None if ent is None else ent[1])
                if all(g is None for g in gouts):
                    continue
                S.active = node.active      # the variables that were trainable when the node was recorded
                gins = node.bw(gouts)
                for inp, gi in zip(node.inputs, gins):
                    if gi is not None and inp.requires_grad:
                        push(inp, gi)
            S.active = prev_active
            for p in waiting:
                S.touched.pop(id(p), None)
                on_ready(p)
            if accumulate is not False and on_ready is None:      # this sweep consumed their gradients all the same
                for k_ in [k_ for k_, p in S.touched.items() if accumulate is True or k_ in accumulate]:
                    del S.touched[k_]
    finally:
        S.accumulate, S.active = prev_acc, prev_active     # an exception in a rule must not corrupt the tape state
    return [grads[id(w)][1] if id(w) in grads else None for w in wrt]


def leaf(buf, shape=None, requires_grad=True):
    t = Tensor(buf, shape)
    t.requires_grad = requires_grad
    return t
