"""The reference's Gen-2 layer API (hem/ops/layers.py:23-356, hem/ops/activations.py, hem/ops/images.py:53-89,
hem/ops/losses.py:10-15) with the same signatures and op order:

    conv / matmul + bias -> [instance norm] -> [batch norm (fused NCHW statistics)] -> activation -> [dropout]

The reference runs these layers in NCHW (`data_format='NCHW'`, hem/ops/layers.py:117-119); activations here are the
engine's NHWC buffers, so its "axis=1" channel concatenations become `engine.concat_channels`, and an NCHW input batch
(session.Input(layout="NCHW"), as the Gen-2 pipeline yields it) is transposed by the first op that touches it
(`rescale`).  `hem.reshape` takes the NHWC shape the reference's own helper takes (hem/ops/layers.py:343-356).
Not on the accelerated path: `use_batch_renorm` (raises).
"""
from .. import _capi as K
from .. import engine as E
from ..ops import layers as _L
from ..ops.activations import make_lrelu
from ..ops.arg_scope import add_arg_scope
from ..variables import xavier_initializer


def _opts(use_batch_renorm, use_instance_norm, dropout, padding, reuse):
    if use_batch_renorm:
        raise K.B200Error("batch renormalisation (renorm=True, hem/ops/layers.py:62,124) is not on the accelerated path")
    if padding not in ('SAME', 'VALID'):
        raise K.B200Error("padding must be 'SAME' or 'VALID'")
    return _L._Opts(instance_norm=bool(use_instance_norm), dropout=dropout or 0, padding=padding, fused_nchw=True,
                    reuse=reuse)


@add_arg_scope
def dense(x, input_size, output_size, init=xavier_initializer, use_batch_norm=False, use_batch_renorm=False,
          activation=None, reuse=False, dropout=0, name=None):
    """hem/ops/layers.py:23-67 (`dropout` is the keep probability, as the reference passes it to tf.nn.dropout)."""
    return _L._dense(x, input_size, output_size, init, use_batch_norm, activation, name,
                     _opts(use_batch_renorm, False, dropout, 'SAME', reuse))


@add_arg_scope
def conv2d(x, input_size, output_size, filter_size=3, stride=1, init=xavier_initializer, use_batch_norm=False,
           use_batch_renorm=False, use_instance_norm=False, activation=None, reuse=False, dropout=0, padding='SAME',
           name=None):
    """hem/ops/layers.py:71-135."""
    return _L._conv2d(x, input_size, output_size, filter_size, stride, init, use_batch_norm, activation, name,
                      _opts(use_batch_renorm, use_instance_norm, dropout, padding, reuse))


@add_arg_scope
def deconv2d(x, input_size, output_size, filter_size=3, stride=2, init=xavier_initializer, output_shape=None,
             use_batch_norm=False, use_batch_renorm=False, use_instance_norm=False, activation=None, reuse=False,
             dropout=0, padding='SAME', name=None):
    """hem/ops/layers.py:139-211.  output_shape: (H, W) of the result (the reference passes the full NCHW shape
    tensor; only its spatial part is free)."""
    if output_shape is not None and len(output_shape) == 4:
        output_shape = tuple(output_shape[2:])                    # [N, C, H, W] as the reference builds it
    return _L._deconv2d(x, input_size, output_size, filter_size, stride, init, use_batch_norm, activation, name,
                        output_shape, _opts(use_batch_renorm, use_instance_norm, dropout, padding, reuse))


@add_arg_scope
def residual(x, input_size, output_size, filter_size=3, stride=1, init=xavier_initializer, use_batch_norm=False,
             use_batch_renorm=False, use_instance_norm=False, activation=None, reuse=False, dropout=0, padding='SAME',
             name=None):
    """hem/ops/layers.py:216-320: two convolutions with a shortcut taken after the first conv + bias."""
    return _L._residual(x, input_size, output_size, filter_size, stride, init, use_batch_norm, activation, name,
                        _opts(use_batch_renorm, use_instance_norm, dropout, padding, reuse))


flatten = _L.flatten


@add_arg_scope
def reshape(x, shape, name=None):
    """hem/ops/layers.py:343-356: `shape` is given in NHWC order (the reference converts it to NCHW; the engine's
    activations already are NHWC)."""
    return E.reshape(x, tuple(shape))


def lrelu(x, leak=0.2, name=None):
    """hem/ops/activations.py:10-28; use `fused_lrelu(leak)` as a layer `activation=` to fuse it."""
    return E.activation(x, K.ACT_LRELU, leak)


fused_lrelu = make_lrelu


def rescale(x, orig=(-1, 1), new=(0, 1), name=None):
    """hem/ops/images.py:53-70: (x - orig0) * (new1-new0)/(orig1-orig0) + new0.  An NCHW / uint8 input batch is
    transposed / normalised in the same pass (engine.affine)."""
    mul = (new[1] - new[0]) / (orig[1] - orig[0])
    return E.affine(x, mul, new[0] - orig[0] * mul)


def instance_norm(x, reuse=False, name=None):
    """hem/ops/images.py:73-89."""
    return _L.instance_norm(x, name)


def to_nchw(x):
    """Hand an activation back in the reference's layout (fp32 [N, C, H, W])."""
    return E.to_nchw(x)


def rmse(x, x_hat, name='rmse'):
    """hem/ops/losses.py:10-11: sqrt(mean((x_hat - x)^2)) -> fp32 [1] holding the MEAN SQUARE; the caller
    takes the square root when it reads the scalar back (reported metric only)."""
    return E.eltloss(x_hat, x, 5, scale=1.0 / x.numel)
