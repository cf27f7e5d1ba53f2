"""The reference's layer API (ops/layers.py:27-166) with identical signatures, NHWC.

conv2d / deconv2d / dense create `<scope>/vars/<name>/weights|bias` on first use and look them up
again when `reuse=True` (ops/layers.py:50-53); op order is conv -> +bias -> batch_norm -> activation
(ops/layers.py:101-105).  Underneath: tcgen05 implicit-GEMM kernels through the C ABI.
"""
import contextlib
import os

from .. import _capi as K
from .. import engine as E
from ..variables import ones_initializer, xavier_initializer, zeros_initializer
from .arg_scope import add_arg_scope

_store = None


def set_store(store):
    global _store
    _store = store


def get_store():
    if _store is None:
        raise K.B200Error("no VariableStore bound (models bind one with ops.layers.set_store)")
    return _store


@contextlib.contextmanager
def variable_scope(name):
    """tf.variable_scope(name) as the models use it (models/gan.py:57-59,201)."""
    st = get_store()
    st.scope.append(name)
    try:
        yield
    finally:
        st.scope.pop()


# Channel counts C with C % 16 == 8 (the reference's 200-channel layers) make every NHWC pixel row start on an
# odd 16-byte boundary, which costs the TMA loads of the GEMM operands about 20%; counts that are not a
# multiple of 8 (the 100-channel deconv of the 64x64 generator at L=200) cannot be addressed by TMA at all.
# Layers wider than 8 channels are therefore stored with their channel count rounded up to a multiple of 16:
# the padded weights / biases are zero, get zero gradients and stay zero, so the TF variables (the leading
# block of each padded array) see exactly the reference's arithmetic.
CHANNEL_PAD = os.environ.get("B200GAN_CHANNEL_PAD", "1") != "0"


def physical_channels(c):
    return (c + 15) // 16 * 16 if (CHANNEL_PAD and c > 8) else c


def _logical_c(x):
    return x.shape[-1] if x.logical_c is None else x.logical_c


def _mark(t, logical_c):
    if t.shape[-1] != logical_c:
        t.logical_c = logical_c
    return t


def weight_name(name):
    return name if name is None else name + '/weights'


def bias_name(name):
    return name if name is None else name + '/bias'


def _variables(name, w_shape, b_shape, init, w_phys=None, b_phys=None):
    st = get_store()
    st.scope.append('vars')
    try:
        W = st.get_variable(weight_name(name), w_shape, init(), w_phys)
        b = st.get_variable(bias_name(name), b_shape, init(), b_phys)
    finally:
        st.scope.pop()
    return W, b


def _fusable(activation):
    if activation is None:
        return K.ACT_NONE, 0.0, True
    code = getattr(activation, 'b200_act', None)
    if code is None:
        return K.ACT_NONE, 0.0, False
    return code[0], code[1], True


def batch_norm(h, activation=None, fused_nchw=False):
    """tf.contrib.layers.batch_norm(h) with defaults (ops/layers.py:10,58): creates
    `<scope>/BatchNorm[_k]/beta` (+ the non-trainable moving_mean / moving_variance, decay 0.999); never shares them
    through `reuse`.  fused_nchw: the Gen-2 call (hem/ops/layers.py:124, fused=True) whose moving
    variance is fed the unbiased batch variance."""
    st = get_store()
    st.scope.append(st.unique_bn_scope())
    try:
        beta = st.get_variable('beta', (_logical_c(h),), zeros_initializer(), (h.shape[-1],))
        mm = st.get_state('moving_mean', (_logical_c(h),), 0.0, (h.shape[-1],))
        mv = st.get_state('moving_variance', (_logical_c(h),), 1.0, (h.shape[-1],))
    finally:
        st.scope.pop()
    act, leak, ok = _fusable(activation)
    out = _mark(E.batch_norm_act(h, beta, act, leak, moving=(mm.buf, mv.buf), unbiased=fused_nchw), _logical_c(h))
    return out if ok else _mark(activation(out), _logical_c(h))


class _Opts:
    """What the Gen-2 layer API adds to a layer (hem/ops/layers.py:23-211); the Gen-1 API uses the defaults."""

    def __init__(self, instance_norm=False, dropout=0, padding='SAME', fused_nchw=False, reuse=False):
        self.instance_norm, self.dropout, self.padding = instance_norm, dropout, padding
        self.fused_nchw, self.reuse = fused_nchw, reuse


_GEN1 = _Opts()


def instance_norm(h, name):
    """hem.instance_norm (hem/ops/images.py:73-89): variables `vars/<name>/shift` (zeros) and `vars/<name>/scale` (ones)."""
    st = get_store()
    c, cp = _logical_c(h), h.shape[-1]
    st.scope.append('vars')
    try:
        shift = st.get_variable(name + '/shift', (c,), zeros_initializer(), (cp,))
        scale = st.get_variable(name + '/scale', (c,), ones_initializer(), (cp,))
    finally:
        st.scope.pop()
    return _mark(E.instance_norm(h, scale, shift), c)


def _finish(h, use_batch_norm, activation, fused, opts=_GEN1, name=None):
    """conv/dense output -> [instance norm] -> [batch norm] -> activation -> [dropout] (hem/ops/layers.py:123-132;
    ops/layers.py:103-104 is the same without the bracketed Gen-2 steps)."""
    if opts.instance_norm:
        h = instance_norm(h, name)
    if use_batch_norm:
        h = batch_norm(h, activation, fused_nchw=opts.fused_nchw)
    elif activation is not None and not fused:
        act, leak, ok = _fusable(activation)
        h = _mark(E.activation(h, act, leak) if ok else activation(h), _logical_c(h))
    if opts.dropout and opts.dropout > 0:
        sess_u = _uniform_like(h)
        h = _mark(E.dropout(h, opts.dropout, sess_u), _logical_c(h))
    return h


def _uniform_like(h):
    from .. import session as S
    return S.current().random_uniform(h.shape, stream_id=2, f32=True)


def _dense(x, input_size, output_size, init, use_batch_norm, activation, name, opts=_GEN1):
    assert _logical_c(x) == input_size, "dense %s: input has %d features, input_size=%d" % (name, _logical_c(x), input_size)
    M = x.shape[0]
    fuse_ok = not (use_batch_norm or opts.instance_norm)
    if output_size == 1:
        if x.logical_c is not None:
            x = E.unpad_channels(x)
        W, b = _variables(name, (input_size, 1), (1,), init)
        act, leak, ok = _fusable(activation)
        fuse = ok and fuse_ok
        h = E.dense_n1(x, W, b, act if fuse else K.ACT_NONE, leak)
        h = E.reshape(h, (M, 1))
    else:
        if x.logical_c is None and input_size % 8:
            x = E.pad_channels(x, (input_size + 15) // 16 * 16)       # e.g. latent_size 50: rows of 16-byte multiples
        cin, cout = x.shape[-1], output_size
        W, b = _variables(name, (input_size, output_size), (output_size,), init, (cin, cout), (cout,))
        act, leak, ok = _fusable(activation)
        fuse = ok and fuse_ok
        g = E.conv_geom(M, 1, 1, cin, cout, 1, 1)
        g.logical = (input_size, output_size)
        xin = E.reshape(x, (M, 1, 1, cin))
        h = E.conv_like('fprop', xin, W, g, bias=b, act=act if fuse else K.ACT_NONE, leak=leak)
        h = _mark(E.reshape(h, (M, cout)), output_size)
    h = _finish(h, use_batch_norm, activation, fuse, opts, name)
    if not opts.reuse:
        get_store().add_to_collection('dense_layers', name, h)
    return h


def _conv_raw(x, input_size, output_size, filter_size, stride, init, name, act, leak, padding):
    """conv + bias (+ a fusable activation): tf.nn.conv2d + bias_add, ops/layers.py:101-102."""
    N, H, Wd, C = x.shape
    assert _logical_c(x) == input_size, "conv2d %s: input has %d channels, input_size=%d" % (name, _logical_c(x), input_size)
    cout = physical_channels(output_size)
    W, b = _variables(name, (filter_size, filter_size, input_size, output_size), (output_size,), init,
                      (filter_size, filter_size, C, cout), (cout,))
    g = E.conv_geom(N, H, Wd, C, cout, filter_size, stride, padding)
    g.logical = (input_size, output_size)
    return _mark(E.conv_like('fprop', x, W, g, bias=b, act=act, leak=leak), output_size)


def _conv2d(x, input_size, output_size, filter_size, stride, init, use_batch_norm, activation, name, opts=_GEN1):
    act, leak, ok = _fusable(activation)
    fuse = ok and not (use_batch_norm or opts.instance_norm)
    h = _conv_raw(x, input_size, output_size, filter_size, stride, init, name, act if fuse else K.ACT_NONE, leak,
                  opts.padding)
    h = _finish(h, use_batch_norm, activation, fuse, opts, name)
    if not opts.reuse:
        get_store().add_to_collection('conv_layers', name, h)
    return h


def _deconv2d(x, input_size, output_size, filter_size, stride, init, use_batch_norm, activation, name, output_shape,
              opts=_GEN1):
    N, h_in, w_in, C = x.shape
    assert _logical_c(x) == input_size, "deconv2d %s: input has %d channels, input_size=%d" % (name, _logical_c(x), input_size)
    cout = physical_channels(output_size)
    W, b = _variables(name, (filter_size, filter_size, output_size, input_size), (output_size,), init,
                      (filter_size, filter_size, cout, C), (cout,))
    act, leak, ok = _fusable(activation)
    fuse = ok and not (use_batch_norm or opts.instance_norm)
    Ho, Wo = (h_in * 2, w_in * 2) if output_shape is None else output_shape
    # the forward conv whose adjoint this is: [N,Ho,Wo,output_size] -> [N,h_in,w_in,input_size]
    g = E.conv_geom(N, Ho, Wo, cout, C, filter_size, stride, opts.padding)
    g.logical = (output_size, input_size)
    if (g.Ho, g.Wo) != (h_in, w_in):
        raise K.B200Error("deconv2d %s: output %dx%d is not a %s stride-%d transpose of the %dx%d input"
                          % (name, Ho, Wo, opts.padding, stride, h_in, w_in))
    h = _mark(E.conv_like('dgrad', x, W, g, bias=b, act=act if fuse else K.ACT_NONE, leak=leak), output_size)
    h = _finish(h, use_batch_norm, activation, fuse, opts, name)
    if not opts.reuse:
        get_store().add_to_collection('conv_layers', name, h)
    return h


def _residual(x, input_size, output_size, filter_size, stride, init, use_batch_norm, activation, name, opts):
    """hem.residual (hem/ops/layers.py:216-320): conv A + bias (= the shortcut) -> norms -> activation -> dropout ->
    conv B + bias -> norms -> + shortcut -> activation -> dropout; variables `<name>A/...`, `<name>B/...`."""
    shortcut = _conv_raw(x, input_size, output_size, filter_size, stride, init, name + 'A', K.ACT_NONE, 0.0, opts.padding)
    h = _finish(shortcut, use_batch_norm, activation, False, opts, name)
    h = _conv_raw(h, output_size, output_size, filter_size, stride, init, name + 'B', K.ACT_NONE, 0.0, opts.padding)
    h = _finish(h, use_batch_norm, None, True, _Opts(opts.instance_norm, 0, opts.padding, opts.fused_nchw), name)
    if h.shape != shortcut.shape:
        raise K.B200Error("residual %s: the two convolutions change the shape (%s vs %s)" % (name, shortcut.shape, h.shape))
    h = _mark(E.add(h, shortcut), output_size)
    return _finish(h, False, activation, False, _Opts(False, opts.dropout, opts.padding, opts.fused_nchw), name)


@add_arg_scope
def dense(x, input_size, output_size, init=xavier_initializer, use_batch_norm=False, activation=None,
          reuse=False, name=None):
    """ops/layers.py:27-62 — act(BN(x W + b)), W [input_size, output_size]."""
    return _dense(x, input_size, output_size, init, use_batch_norm, activation, name, _Opts(reuse=reuse))


@add_arg_scope
def conv2d(x, input_size, output_size, filter_size=3, stride=1, init=xavier_initializer, use_batch_norm=False,
           activation=None, reuse=False, name=None):
    """ops/layers.py:66-107 — act(BN(conv_SAME(x, K) + b)), K [k,k,input_size,output_size]."""
    return _conv2d(x, input_size, output_size, filter_size, stride, init, use_batch_norm, activation, name,
                   _Opts(reuse=reuse))


@add_arg_scope
def deconv2d(x, input_size, output_size, filter_size=3, stride=2, init=xavier_initializer, use_batch_norm=False,
             activation=None, reuse=False, name=None, output_shape=None):
    """ops/layers.py:111-148 — act(BN(conv2d_transpose_SAME(x, K) + b)), K [k,k,output_size,input_size];
    output is 2x the input (ops/layers.py:141) unless output_shape=(H,W) is given (the Gen-2 kwarg,
    hem/ops/layers.py:185-187, used by the shape-generalised autoencoders)."""
    return _deconv2d(x, input_size, output_size, filter_size, stride, init, use_batch_norm, activation, name,
                     output_shape, _Opts(reuse=reuse))


@add_arg_scope
def flatten(x, name=None):
    """ops/layers.py:152-166 — [B, -1]."""
    return E.reshape(x, (x.shape[0], -1))          # (engine.reshape strips zero channel padding first)
