"""Summaries after the step — the reference's ops/summaries.py (summarize_activations 13-23, summarize_losses 26-30,
summarize_weights_biases 33-44, summarize_gradients 47-57, factorization 83-96, montage_summary 99-127) as cheap device
reductions through the C ABI (b200_summary_stats, b200_montage).

The reference attaches tf.summary ops to the graph and evaluates them with extra `sess.run(summary_op)` calls
(train.py:291,311-316,327); here a summary pass is an extra forward run under `collecting(store)`, during which the layers
register their outputs in the 'conv_layers' / 'dense_layers' collections (ops/layers.py:60,105,146), and the
`summarize_*` functions return plain dicts: {name: {"histogram": {...}, "sparsity": float[, "montage": tensor]}}.
Writing TensorBoard event files is outside the hot path.
"""
import contextlib
import math

import torch

from .. import engine as E

N_BUCKETS = 1550          # 774 exponential buckets (1e-12 * 1.1^k) per sign + the two around zero, as TensorBoard's default


@contextlib.contextmanager
def collecting(store):
    """Layers called inside register their outputs (tf.add_to_collection in the reference's layers)."""
    prev, store.collect = store.collect, True
    store.collections = {}
    try:
        yield store.collections
    finally:
        store.collect = prev


def tensor_stats(t, buckets=True):
    """tf.summary.histogram(t) + tf.nn.zero_fraction(t) in one pass over the device buffer."""
    buf = t.buf if isinstance(t, E.Tensor) else t
    buf = buf.contiguous()
    kind = 1 if buf.dtype == torch.float32 else 0
    if buf.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("summaries take fp32 / bf16 device tensors")
    out5 = torch.empty(5, dtype=torch.float32, device=buf.device)
    counts = torch.empty(N_BUCKETS, dtype=torch.int32, device=buf.device) if buckets else None
    E.launch("b200_summary_stats", E._p(buf), kind, buf.numel(), E._p(out5), E._p(counts), N_BUCKETS if buckets else 0, n=3)
    mn, mx, s, q, z = [float(v) for v in out5.tolist()]
    n = buf.numel()
    res = {"histogram": {"min": mn, "max": mx, "num": n, "sum": s, "sum_squares": q}, "sparsity": z / n}
    if buckets:
        res["histogram"]["bucket_counts"] = counts.cpu()
    return res


def bucket_edges():
    """Lower edge magnitude of bucket k >= 1 on either side of zero: 1e-12 * 1.1^(k-1); bucket 0 is |v| < 1e-12."""
    half = N_BUCKETS // 2
    return [0.0] + [1e-12 * 1.1 ** (k - 1) for k in range(1, half)]


def factorization(n):
    """ops/summaries.py:83-96: (rows, cols) with rows <= sqrt(n) dividing n."""
    for i in range(int(math.sqrt(float(n))), 0, -1):
        if n % i == 0:
            return (i, int(n / i))


def montage_summary(x, m=0, n=0, name=None, rescale=(1.0, 0.0)):
    """ops/summaries.py:99-127: [m*n, H, W(, C)] images -> one [m*H, n*W, C] fp32 grid (image j at row j % m,
    column j // m).  `x` is an NHWC engine Tensor / device tensor; `rescale` = (mul, add) applied on the way."""
    buf = x.torch() if isinstance(x, E.Tensor) else x
    if buf.dim() == 3:
        buf = buf.unsqueeze(-1)
    count, h, w, c = buf.shape
    if n == 0 or m == 0:
        m, n = factorization(count)
    if m * n > count:
        raise ValueError("montage of %d x %d needs %d images, got %d" % (m, n, m * n, count))
    buf = buf.contiguous()
    out = torch.empty((m * h, n * w, c), dtype=torch.float32, device=buf.device)
    E.launch("b200_montage", E._p(buf), 1 if buf.dtype == torch.float32 else 0, E._p(out), m, n, h, w, c,
             float(rescale[0]), float(rescale[1]))
    return out


def summarize_activations(store):
    """ops/summaries.py:13-23: histogram + sparsity of every layer output, and for conv layers a montage of the
    first example's channels (`tf.transpose(l[0], [2, 0, 1])`)."""
    out = {}
    for name, t in store.collections.get('conv_layers', []):
        ent = tensor_stats(t)
        first = t.torch()[0]                                        # [H, W, C] of example 0
        c = first.shape[-1] if t.logical_c is None else t.logical_c
        ent["montage"] = montage_summary(first[..., :c].permute(2, 0, 1).contiguous())
        out["activations/" + name] = ent
    for name, t in store.collections.get('dense_layers', []):
        out["activations/" + name] = tensor_stats(t)
    return out


def summarize_losses(losses):
    """ops/summaries.py:26-30: {name: fp32 [1] device tensor} -> scalars."""
    return {"loss/" + k: float(v.item() if hasattr(v, "item") else v.buf.item()) for k, v in losses.items()}


def summarize_weights_biases(store):
    """ops/summaries.py:33-44: histogram + sparsity of every weight and bias (TF-named, logical shapes)."""
    out = {}
    for name, p in store.params.items():
        kind = "weights" if name.endswith("/weights") else ("biases" if name.endswith("/bias") else None)
        if kind:
            out["%s/%s" % (kind, name)] = tensor_stats(p.logical(p.p32).contiguous())
    return out


def summarize_gradients(store, name='gradients'):
    """ops/summaries.py:47-57: histogram of every variable's (averaged) gradient currently in the buckets."""
    return {"%s/%s/gradient" % (name, n): tensor_stats(p.logical(p.g32).contiguous()) for n, p in store.params.items()}
