"""ops/input.py:11-25 — tower i takes rows [i*B, (i+1)*B) of the global batch."""
from .. import engine as E


def batch_slice(x, batch_size, slice_index, name=None):
    t = x.torch()[slice_index * batch_size:(slice_index + 1) * batch_size]
    return E.Tensor(t.contiguous() if not t.is_contiguous() else t)
