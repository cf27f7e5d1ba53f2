"""Convolutional autoencoder on the B200 engine — the reference's models/cnn.py (cnn 20-59, loss 75-79,
latent 82-93, encoder 96-112, decoder 115-134) with the shape generalisation of SURVEY App. C #1: the
bottleneck side s = the image side after four k5 s2 convs (4 at 64x64) and the decoder's deconvs mirror the
encoder's sizes through `output_shape` (so 28x28x1 MNIST-shaped inputs work: 28->14->7->4->2 and back)."""
from .. import engine as E
from .. import session as S
from ..ops.activations import lrelu, relu, tanh
from ..ops.arg_scope import arg_scope
from ..ops.layers import conv2d, deconv2d, dense, flatten, variable_scope
from ..variables import optimizer_cfg


def encoder_sizes(H):
    sizes = [H]
    for _ in range(4):
        sizes.append(-(-sizes[-1] // 2))
    return sizes


def cnn(x, args):
    """models/cnn.py:20-59; returns default_training's helper (util.py:22-28)."""
    sess = S.current()
    store = sess.store
    H, W, C = x.shape
    sizes = encoder_sizes(H)

    def tower(batch01):
        with E.recording(True, active='all'):
            with variable_scope('rescale'):
                xr = E.affine(batch01, 2.0, -1.0)                     # x = 2*(x-0.5)  cnn.py:30-31
            with variable_scope('encoder'):
                e = encoder(xr, C)
            with variable_scope('latent'):
                z = latent(e, args.latent_size, sizes[-1])
            with variable_scope('decoder'):
                d = decoder(z, args.latent_size, sizes, C)
            with variable_scope('loss'):
                d_loss = loss(xr, d)
        return d_loss

    return _default_training(sess, x, args, tower, lambda out: {'loss': out})


def _default_training(sess, x, args, tower, name_losses):
    """util.py:22-28 + opt.compute_gradients / average_gradients / apply_gradients (cnn.py:42,50-52)."""
    store = sess.store
    with sess.building():                                             # graph construction: variables only
        store.begin_pass()
        out = tower(x.next())
        E.backward([((out[0] if isinstance(out, tuple) else out), None)])
    params = list(store.params.values())
    group = None
    if sess.cuda:
        x.materialize(sess.device)
        store.finalize([('all', params, optimizer_cfg(args))], sess.device)
        (group,) = store.groups

    ex = sess.exchange(group) if sess.cuda else None

    def iteration():
        x.reset()                                     # (the gradient bucket is zero: apply_gradients resets it)
        out = tower(x.next())
        ex.begin()
        E.backward([((out[0] if isinstance(out, tuple) else out), None)], on_ready=ex.on_ready)
        ex.finish()
        group.apply_gradients(ex.join(), 0.0)
        return {k: v.buf for k, v in name_losses(out).items()}

    def helper(sess_, args_):
        res = sess.run('ae_iteration', iteration)
        return {k: float(v.item()) for k, v in res.items()}

    helper.iteration = iteration
    helper.tower = tower
    helper.store = store
    return helper


def loss(x, d):
    """mean(|x - d|) — cnn.py:75-79."""
    return E.eltloss(d, x, 0, scale=1.0 / d.numel)


def latent(x, latent_size, s, reuse=False):
    with arg_scope([dense], reuse=reuse):
        x = flatten(x)
        x = dense(x, 32 * s * s, latent_size, name='d1')
    return x


def encoder(x, C=3, reuse=False):
    with arg_scope([conv2d], reuse=reuse, activation=lrelu):
        x = conv2d(x, C, 64, 5, 2, name='c1')
        x = conv2d(x, 64, 128, 5, 2, name='c2')
        x = conv2d(x, 128, 256, 5, 2, name='c3')
        x = conv2d(x, 256, 256, 5, 2, name='c4')
        x = conv2d(x, 256, 96, 1, name='c5')
        x = conv2d(x, 96, 32, 1, name='c6')
    return x


def decoder(x, latent_size, sizes, C=3, reuse=False, final=tanh):
    s = sizes[-1]
    with arg_scope([dense, conv2d, deconv2d], reuse=reuse, activation=relu):
        x = dense(x, latent_size, 32 * s * s, name='d1')
        x = E.reshape(x, (-1, s, s, 32))                              # un-flatten
        x = conv2d(x, 32, 96, 1, name='c1')
        x = conv2d(x, 96, 256, 1, name='c2')
        x = deconv2d(x, 256, 256, 5, 2, name='dc1', output_shape=(sizes[3], sizes[3]))
        x = deconv2d(x, 256, 128, 5, 2, name='dc2', output_shape=(sizes[2], sizes[2]))
        x = deconv2d(x, 128, 64, 5, 2, name='dc3', output_shape=(sizes[1], sizes[1]))
        with E.f32_outputs():       # the reconstruction feeds the loss directly: keep it in fp32
            x = deconv2d(x, 64, C, 5, 2, name='dc4', activation=final, output_shape=(sizes[0], sizes[0]))
    return x
