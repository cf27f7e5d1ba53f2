"""GAN / WGAN / Improved WGAN on the B200 engine — same structure and names as the reference's
models/gan.py (gan 39-91, _train_* 110-175, losses 178-211, gradient_penalty 214-231, generator
234-254, discriminator 257-287), with the minimal shape generalisation of SURVEY App. C #1:
the image side H (=W, 4*2^n) and channel count C come from the input instead of the hard-wired
64x64x3; at 64x64x3 this is the reference layer for layer.

One process drives one GPU (= one tower of util.py:54-77): the tower loop is replaced by ranks,
`average_gradients` by an NCCL all-reduce of the flat gradient bucket.
"""
import math

from .. import _capi as K
from .. import engine as E
from .. import session as S
from ..ops.activations import lrelu, relu, sigmoid, tanh
from ..ops.arg_scope import arg_scope
from ..ops.layers import conv2d, deconv2d, dense, flatten, variable_scope
from ..variables import optimizer_cfg


def gan(x, args):
    """Initialize model; returns train_func(sess, args) -> {loss_name: float} (train.py:237-246,307)."""
    sess = S.current()
    store = sess.store
    B = args.batch_size
    H, W, C = x.shape
    assert H == W, "GAN family expects square images"

    def tower(batch01, train, before_critic=None):
        """One sess.run-equivalent on this rank's tower: forward + losses.  `train` in {'d','g'}
        selects whose variables are being differentiated (models/gan.py:55-68).  `before_critic` runs
        between the generator forward and the first critic kernel (the previous run's critic update, which
        overlaps the generator forward, is joined there)."""
        g_params = store.collection('generator')
        d_params = store.collection('discriminator')
        active = {'d': d_params, 'g': g_params, 'dg': d_params + g_params}[train]
        xr = flatten(E.affine(batch01, 2.0, -1.0))                    # x = 2*(x-0.5)  gan.py:49-50
        with E.recording(True, active=active):
            with variable_scope('generator'), E.recording(train != 'd'):   # G is a constant for d_loss
                g = generator(B, args.latent_size, args, H, C)
            if before_critic is not None:
                before_critic()
            with variable_scope('discriminator'):
                d_real = discriminator(xr, args, H, C)
                d_fake = discriminator(g, args, H, C, reuse=True)
            g_loss, d_loss = losses(xr, g, d_fake, d_real, args, H, C, second_order=('d' in train))
        return g_loss, d_loss

    with sess.building():                                             # graph construction: variables only
        for mode in ('d', 'g'):
            store.begin_pass()
            gl, dl = tower(x.next(), mode)
            E.backward([(dl if mode == 'd' else gl, None)])
    g_params = store.collection('generator')
    d_params = store.collection('discriminator')
    if sess.cuda:
        x.materialize(sess.device)
        store.finalize([('generator', g_params, optimizer_cfg(args)),
                        ('discriminator', d_params, optimizer_cfg(args))], sess.device)
        g_group, d_group = store.groups
    else:
        g_group = d_group = None

    clip = 0.01 if args.model == 'wgan' else 0.0                      # gan.py:142-143

    pending = [False]                             # a critic exchange is in flight on the side stream
    ex_d = sess.exchange(d_group) if sess.cuda else None
    ex_g = sess.exchange(g_group) if sess.cuda else None

    def finish_pending():
        """Join the deferred gradient exchange of the previous critic run and apply its update."""
        if pending[0]:
            d_group.apply_gradients(ex_d.join(), clip)
            pending[0] = False

    def before_critic_d():
        finish_pending()          # (every gradient bucket is zero here: apply_gradients resets it in the same pass)

    def d_run():
        E.S.bn_updates = True                     # d_train_op depends on batchnorm_updates in all three schedules
        g_loss, d_loss = tower(x.next(), 'd', before_critic_d)
        E.S.bn_updates = False
        # average_gradients: buckets of the critic's gradient go out as the sweep finishes them (c3 + fc2 first)
        ex_d.begin()
        E.backward([(d_loss, None)], on_ready=ex_d.on_ready)
        ex_d.finish()
        if ex_d.active and sess.overlap_updates:
            # multi-GPU: the tail of the exchange also overlaps the next run's generator forward (which reads no
            # critic variable); the update itself is applied when that run reaches the critic
            pending[0] = True
        else:
            d_group.apply_gradients(ex_d.join(), clip)
        return g_loss, d_loss

    def g_run():
        E.S.bn_updates = args.model == 'wgan'     # gan.py:145-148 (wgan) vs 163 (iwgan: g_train_op has no dependency)
        g_loss, d_loss = tower(x.next(), 'g', finish_pending)
        E.S.bn_updates = False
        ex_g.begin()
        E.backward([(g_loss, None)], on_ready=ex_g.on_ready)
        ex_g.finish()
        g_group.apply_gradients(ex_g.join(), clip)
        return g_loss, d_loss

    def gan_run():                                                    # _train_gan: one run, both updates
        E.S.bn_updates = True                                         # gan.py:126-128
        gl, dl = tower(x.next(), 'dg')                                # same forward for both (App. C #7)
        E.S.bn_updates = False
        ex_d.begin(); ex_g.begin()
        E.backward([(dl, None)], accumulate=store.collection('discriminator'), on_ready=ex_d.on_ready)
        ex_d.finish()
        E.backward([(gl, None)], accumulate=store.collection('generator'), on_ready=ex_g.on_ready)
        ex_g.finish()
        d_group.apply_gradients(ex_d.join(), 0.0)
        g_group.apply_gradients(ex_g.join(), 0.0)
        return gl, dl

    def iteration():
        x.reset()
        if args.model == 'gan':
            gl, dl = gan_run()
        else:
            for _ in range(args.n_disc_train):                        # gan.py:152-153,170-171
                d_run()
                store.begin_pass()
            gl, dl = g_run()
        finish_pending()
        return {'g_loss': gl.buf, 'd_loss': dl.buf}

    def helper(sess_, args_):
        out = sess.run('gan_iteration', iteration)
        return {k: float(v.item()) for k, v in out.items()}

    helper.iteration = iteration
    helper.store = store
    helper.tower = tower
    return helper


def losses(x, g, d_fake, d_real, args, H, C, second_order=True):
    """models/gan.py:178-211.  Returns (g_loss, d_loss) as fp32 [1] device tensors."""
    if args.model == 'gan':
        n = d_fake.shape[0]
        g_loss = E.eltloss(d_fake, None, 1, scale=1.0 / n)                       # mean(-log(d_fake+1e-8))
        d_loss = E.add_scalars(E.eltloss(d_real, None, 1, scale=1.0 / n),
                               E.eltloss(d_fake, None, 2, scale=1.0 / n))
        return g_loss, d_loss
    if args.model == 'wgan':
        return E.wgan_losses(d_real, d_fake)
    with variable_scope('discriminator'):                             # gan.py:201
        ss = gradient_penalty(x, g, args, H, C, second_order)
    return E.wgan_losses(d_real, d_fake, ss, 10.0)


def gradient_penalty(x, g, args, H, C, second_order=True):
    """models/gan.py:214-231: separate critic pass on x + alpha (g - x); ONE norm over the whole
    tower batch (App. C #3).  Returns sum(gradients^2); sqrt / (.-1)^2 / lambda live in the fused
    loss kernel."""
    sess = S.current()
    B = args.batch_size
    alpha = sess.random_uniform((B, 1))
    interpolates = E.interpolate(x, g, alpha)
    interpolates.requires_grad = True
    interpolates.grad_f32 = True
    d_interpolates = discriminator(interpolates, args, H, C, reuse=True)
    ones = E.Tensor(E.empty(d_interpolates.shape, E.F32))
    E.launch("b200_fill_f32", E._p(ones.buf), ones.numel, 1.0)
    (gradients,) = E.backward([(d_interpolates, ones)], wrt=[interpolates], create_graph=second_order,
                              accumulate=False)                       # tf.gradients(d_interpolates, [interpolates])
    return E.sumsq(gradients)


def n_up_stages(H):
    n = int(round(math.log2(H / 4)))
    assert 4 * 2 ** n == H, "generator needs H = 4*2^n"
    return n


def generator(batch_size, latent_size, args, H=64, C=3, reuse=False):
    """models/gan.py:234-254: z -> fc1 -> [4,4,4L] -> n deconvs (k5 s2, BN+relu; last: tanh, no BN)."""
    sess = S.current()
    n = n_up_stages(H)
    with arg_scope([dense, deconv2d], reuse=reuse, use_batch_norm=True, activation=relu):
        z = sess.random_normal((batch_size, latent_size))
        y = dense(z, latent_size, 4 * 4 * 4 * latent_size, name='fc1')
        y = E.reshape(y, (-1, 4, 4, 4 * latent_size))
        cin = 4 * latent_size
        for i in range(1, n + 1):
            if i < n:
                y = deconv2d(y, cin, cin // 2, 5, 2, name='dc%d' % i)
                cin //= 2
            else:
                y = deconv2d(y, cin, C, 5, 2, name='dc%d' % i, activation=tanh, use_batch_norm=False)
        y = E.reshape(y, (-1, H * H * C))
    return y


def discriminator(x, args, H=64, C=3, reuse=False):
    """models/gan.py:257-287: 3 convs k5 s2 + lrelu (BN on c2,c3 unless iwgan), reshape to rows of
    4*4*4L (4 rows/image at 64x64, App. C #2), dense -> 1 (sigmoid for gan)."""
    use_bn = False if args.model == 'iwgan' else True
    final_activation = None if args.model in ['wgan', 'iwgan'] else sigmoid
    L = args.latent_size
    with arg_scope([conv2d], use_batch_norm=use_bn, activation=lrelu, reuse=reuse):
        x = E.reshape(x, (-1, H, H, C))
        x = conv2d(x, C, L, 5, 2, name='c1', use_batch_norm=False)
        x = conv2d(x, L, L * 2, 5, 2, name='c2')
        x = conv2d(x, L * 2, L * 4, 5, 2, name='c3')
        x = E.reshape(x, (-1, 4 * 4 * 4 * L))
        x = dense(x, 4 * 4 * 4 * L, 1, use_batch_norm=False, activation=final_activation, name='fc2', reuse=reuse)
        x = E.reshape(x, (-1,))
    return x
