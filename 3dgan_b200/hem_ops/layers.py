"""The reference's Gen-2 layer API (hem/ops/layers.py:23-356, hem/ops/activations.py, hem/ops/images.py:53-70,
hem/ops/losses.py:10-15) with the same signatures.

The reference runs these layers in NCHW (`data_format='NCHW'`, hem/ops/layers.py:117-119); the tensors here
are the engine's NHWC buffers, so "axis=1" channel concatenations of the reference become
`engine.concat_channels`.  Math is identical (fused batch-norm = same batch statistics, beta only).
Kwargs that no in-scope model uses (batch-renorm, instance-norm, dropout, VALID padding, `residual`) raise.
"""
from .. import _capi as K
from .. import engine as E
from ..ops import layers as _L
from ..ops.activations import make_lrelu
from ..ops.arg_scope import add_arg_scope
from ..variables import xavier_initializer


def _unsupported(use_batch_renorm, use_instance_norm, dropout, padding):
    if use_batch_renorm or use_instance_norm:
        raise K.B200Error("batch-renorm / instance-norm are outside the accelerated path (SURVEY §2 row 8)")
    if dropout and dropout > 0:
        raise K.B200Error("dropout is outside the accelerated path (default 0, hem/models/pix2pix.py:49-52)")
    if padding != 'SAME':
        raise K.B200Error("only SAME padding is on the accelerated path")


@add_arg_scope
def dense(x, input_size, output_size, init=xavier_initializer, use_batch_norm=False, use_batch_renorm=False,
          activation=None, reuse=False, dropout=0, name=None):
    """hem/ops/layers.py:23-67."""
    _unsupported(use_batch_renorm, False, dropout, 'SAME')
    return _L.dense._arg_scope_target(x, input_size, output_size, init=init, use_batch_norm=use_batch_norm,
                                      activation=activation, reuse=reuse, name=name)


@add_arg_scope
def conv2d(x, input_size, output_size, filter_size=3, stride=1, init=xavier_initializer, use_batch_norm=False,
           use_batch_renorm=False, use_instance_norm=False, activation=None, reuse=False, dropout=0, padding='SAME',
           name=None):
    """hem/ops/layers.py:71-135."""
    _unsupported(use_batch_renorm, use_instance_norm, dropout, padding)
    return _L.conv2d._arg_scope_target(x, input_size, output_size, filter_size, stride, init=init,
                                       use_batch_norm=use_batch_norm, activation=activation, reuse=reuse, name=name)


@add_arg_scope
def deconv2d(x, input_size, output_size, filter_size=3, stride=2, init=xavier_initializer, output_shape=None,
             use_batch_norm=False, use_batch_renorm=False, use_instance_norm=False, activation=None, reuse=False,
             dropout=0, padding='SAME', name=None):
    """hem/ops/layers.py:139-211."""
    _unsupported(use_batch_renorm, use_instance_norm, dropout, padding)
    return _L.deconv2d._arg_scope_target(x, input_size, output_size, filter_size, stride, init=init,
                                         use_batch_norm=use_batch_norm, activation=activation, reuse=reuse, name=name,
                                         output_shape=output_shape)


flatten = _L.flatten


def lrelu(x, leak=0.2, name=None):
    """hem/ops/activations.py:10-28; use `fused_lrelu(leak)` as a layer `activation=` to fuse it."""
    return E.activation(x, K.ACT_LRELU, leak)


fused_lrelu = make_lrelu


def rescale(x, orig=(-1, 1), new=(0, 1), name=None):
    """hem/ops/images.py:53-70: (x - orig0) * (new1-new0)/(orig1-orig0) + new0."""
    mul = (new[1] - new[0]) / (orig[1] - orig[0])
    return E.affine(x, mul, new[0] - orig[0] * mul)


def rmse(x, x_hat, name='rmse'):
    """hem/ops/losses.py:10-11: sqrt(mean((x_hat - x)^2)) -> fp32 [1] holding the MEAN SQUARE; the caller
    takes the square root when it reads the scalar back (reported metric only)."""
    return E.eltloss(x_hat, x, 5, scale=1.0 / x.numel)
