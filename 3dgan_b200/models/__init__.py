"""Model builders of the hot path: name -> (builder(x, args) -> train_func, input slots per iteration)."""
from . import cnn as _cnn
from . import gan as _gan
from . import vae as _vae
from .pix2pix import pix2pix  # noqa: F401  (Gen-2 plugin class: pix2pix(x_y, args).train(sess, args, feed))

MODEL_FUNCS = {
    'gan': (_gan.gan, lambda a: 1),
    'wgan': (_gan.gan, lambda a: a.n_disc_train + 1),
    'iwgan': (_gan.gan, lambda a: a.n_disc_train + 1),
    'cnn': (_cnn.cnn, lambda a: 1),
    'vae': (_vae.vae, lambda a: 1),
}
