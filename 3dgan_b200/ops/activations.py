"""Activation functions with the reference's signatures (ops/activations.py:11-29).

Each callable carries `b200_act = (code, leak)` so that `conv2d / deconv2d / dense` can fuse it
into the producing kernel's epilogue; calling it directly runs the stand-alone kernel.
"""
from .. import _capi as K
from .. import engine as E


def lrelu(x, leak=0.2, name=None):
    """max(leak*x, x) — ops/activations.py:28; gradient slope = leak for x <= 0."""
    return E.activation(x, K.ACT_LRELU, leak)


def relu(x, name=None):
    return E.activation(x, K.ACT_RELU, 0.0)


def tanh(x, name=None):
    return E.activation(x, K.ACT_TANH, 0.0)


def sigmoid(x, name=None):
    return E.activation(x, K.ACT_SIGMOID, 0.0)


lrelu.b200_act = (K.ACT_LRELU, 0.2)
relu.b200_act = (K.ACT_RELU, 0.0)
tanh.b200_act = (K.ACT_TANH, 0.0)
sigmoid.b200_act = (K.ACT_SIGMOID, 0.0)


def make_lrelu(leak):
    """An lrelu with a different leak that still fuses (hem.lrelu(x, leak=0), pix2pix.py:202)."""
    def f(x, name=None):
        return E.activation(x, K.ACT_LRELU, leak)
    f.b200_act = (K.ACT_LRELU, float(leak))
    return f
