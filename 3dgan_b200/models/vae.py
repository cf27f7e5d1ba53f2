"""Variational autoencoder on the B200 engine — the reference's models/vae.py (vae 25-51, losses 66-90,
encoder 93-110, latent 113-129, decoder 132-151), shape-generalised like models/cnn.py.

Reference quirks kept: x stays in [0,1]; sigma is the raw dense output; losses are SUMS
over the batch; only `decoder_loss` is differentiated (vae.py:41) — KL and total are reported.  The second
decoder pass on pure noise (`d_fake`, vae.py:37) feeds summaries only and is not part of
sess.run([train_op, losses]), so it is not executed."""
from .. import engine as E
from .. import session as S
from ..ops.activations import lrelu, sigmoid
from ..ops.arg_scope import arg_scope
from ..ops.layers import conv2d, dense, flatten, variable_scope
from . import cnn as _cnn


def vae(x, args):
    sess = S.current()
    store = sess.store
    H, W, C = x.shape
    sizes = _cnn.encoder_sizes(H)

    def tower(batch01):
        with E.recording(True, active='all'):
            xb = E.affine(batch01, 1.0, 0.0)                          # [0,1] input, bf16 copy
            with variable_scope('encoder'):
                e = encoder(xb, C)
            with variable_scope('latent'):
                samples, z, z_mean, z_stddev = latent(e, args.batch_size, args.latent_size, sizes[-1])
            with variable_scope('decoder'):
                d_real = _cnn.decoder(z, args.latent_size, sizes, C, final=sigmoid)
            d_loss, l_loss, t_loss = losses(xb, z_mean, z_stddev, d_real)
        return d_loss, l_loss, t_loss

    return _cnn._default_training(sess, x, args, tower,
                                  lambda out: {'decoder_loss': out[0], 'latent_loss': out[1], 'total_loss': out[2]})


def losses(x, z_mean, z_stddev, d_real):
    """vae.py:66-90: Bernoulli reconstruction (sum), KL (sum), total."""
    d_loss = E.eltloss(d_real, x, 3, scale=1.0)
    with E.recording(False):                                          # reported only (vae.py:41)
        l_loss = E.add_scalars(E.eltloss(z_mean, None, 6), E.eltloss(z_stddev, None, 7))
        t_loss = E.add_scalars(d_loss, l_loss)
    return d_loss, l_loss, t_loss


def encoder(x, C=3, reuse=False):
    with arg_scope([conv2d], reuse=reuse, activation=lrelu, use_batch_norm=True):
        x = conv2d(x, C, 64, 5, 2, name='c1')
        x = conv2d(x, 64, 128, 5, 2, name='c2')
        x = conv2d(x, 128, 256, 5, 2, name='c3')
        x = conv2d(x, 256, 256, 5, 2, name='c4')
        x = conv2d(x, 256, 96, 1, name='c5')
        x = conv2d(x, 96, 32, 1, name='c6')
    return x


def latent(x, batch_size, latent_size, s, reuse=False):
    sess = S.current()
    with arg_scope([dense], reuse=reuse):
        flat = flatten(x)
        z_mean = dense(flat, 32 * s * s, latent_size, name='d1')
        z_stddev = dense(flat, 32 * s * s, latent_size, name='d2')
        samples = sess.random_normal((batch_size, latent_size))
        z = E.reparameterize(z_mean, z_stddev, samples)
    return (samples, z, z_mean, z_stddev)
