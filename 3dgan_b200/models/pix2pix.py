"""pix2pix conditional GAN on the B200 engine — port of the reference's hem/models/pix2pix.py
(__init__ 81-148, train 151-156, generator 160-228, discriminator 232-259, loss 263-304).

U-Net generator (8 x conv k4 s2 down to 1x1, 8 x deconv k4 s2 up with skip concatenations), PatchGAN
discriminator on concat(rgb, depth) -> [B,8,8,1] logits, sigmoid cross-entropy losses (+10*L1 with --add_l1).
Reference quirks kept: skips are unconditional, decoder batch-norm also covers the tanh
output layer d8, weights AND biases are N(0, 0.02), every sess.run (incl. the losses-only run) consumes a
fresh batch.  Tensors are NHWC here (the reference is NCHW: its axis-1 concat is our channel concat).
"""
import math

from .. import _capi as K
from .. import engine as E
from .. import session as S
from ..hem_ops import layers as hem
from ..ops.activations import tanh
from ..ops.arg_scope import arg_scope
from ..ops.layers import variable_scope
from ..variables import optimizer_cfg, random_normal_initializer


class pix2pix:
    name = 'pix2pix'

    @staticmethod
    def arguments():
        """hem/models/pix2pix.py:36-77 (the flags the accelerated path honours)."""
        return {
            '--batch_norm_disc': {'action': 'store_true', 'default': False},
            '--batch_norm_gen': {'action': 'store_true', 'default': False},
            '--n_disc_train': {'type': int, 'default': 1},
            '--add_l1': {'action': 'store_true', 'default': False},
            '--dropout': {'type': float, 'default': 0},
            '--noise': {'type': str, 'nargs': '*', 'default': []},
        }

    def __init__(self, x_y, args):
        """x_y: (rgb Input [B,H,W,3], depth Input [B,H,W,1]) in [0,1]."""
        sess = S.current()
        store = sess.store
        self.sess, self.args = sess, args
        self.x_in, self.y_in = x_y
        if getattr(args, 'noise', None):
            raise K.B200Error("--noise injection is outside the accelerated path")

        def tower(bx, by, train):
            g_params = store.collection('generator')
            d_params = store.collection('discriminator')
            active = {'d': d_params, 'g': g_params, 'none': []}[train]
            x = hem.rescale(bx, (0, 1), (-1, 1))
            y = hem.rescale(by, (0, 1), (-1, 1))
            with E.recording(True, active=active):
                with variable_scope('generator'), E.recording(train == 'g'):
                    g = pix2pix.generator(x, args)
                with variable_scope('discriminator'):
                    d_real_logits = pix2pix.discriminator(x, y, args)
                    d_fake_logits = pix2pix.discriminator(x, g, args, reuse=True)
                return pix2pix.loss(d_real_logits, d_fake_logits, g, y, args)

        self.tower = tower
        with sess.building():
            for mode in ('d', 'g'):
                store.begin_pass()
                ls = tower(self.x_in.next(), self.y_in.next(), mode)
                E.backward([(ls['d_total'] if mode == 'd' else ls['g_total'], None)])
        if sess.cuda:
            self.x_in.materialize(sess.device)
            self.y_in.materialize(sess.device)
            store.finalize([('generator', store.collection('generator'), optimizer_cfg(args)),
                            ('discriminator', store.collection('discriminator'), optimizer_cfg(args))], sess.device)
            self.g_group, self.d_group = store.groups

    # ------------------------------------------------------------------ training (pix2pix.py:151-156)
    def _run(self, mode):
        grp = {'d': self.d_group, 'g': self.g_group}.get(mode)      # (its gradient bucket is zero: see apply_gradients)
        E.S.bn_updates = grp is not None               # d_train_op / g_train_op depend on batchnorm_updates (pix2pix.py:145-147)
        ls = self.tower(self.x_in.next(), self.y_in.next(), mode)
        E.S.bn_updates = False
        if grp is not None:
            ex = self.sess.exchange(grp)                 # buckets go out under the rest of the backward (229 MB/step)
            ex.begin()
            E.backward([(ls['d_total'] if mode == 'd' else ls['g_total'], None)], on_ready=ex.on_ready)
            ex.finish()
            grp.apply_gradients(ex.join(), 0.0)
        return ls

    def iteration(self):
        self.x_in.reset(); self.y_in.reset()
        for _ in range(self.args.n_disc_train):
            self._run('d')
            self.sess.store.begin_pass()
        self._run('g')
        self.sess.store.begin_pass()
        ls = self._run('none')                         # sess.run(self.all_losses): forward only, fresh batch
        return {k: v.buf for k, v in ls.items()}

    def train(self, sess, args, feed_dict=None):
        out = self.sess.run('pix2pix_iteration', self.iteration)
        res = {k: float(v.item()) for k, v in out.items()}
        res['rmse'] = math.sqrt(max(res['rmse'], 0.0))
        return res

    # ------------------------------------------------------------------ graph
    @staticmethod
    def generator(x, args, reuse=False):
        init = lambda: random_normal_initializer(mean=0, stddev=0.02)
        with arg_scope([hem.conv2d], reuse=reuse, use_batch_norm=args.batch_norm_gen, filter_size=4, stride=2,
                       init=init, activation=hem.fused_lrelu(0.2)):
            with variable_scope('enocder'):                      # sic (pix2pix.py:182)
                e1 = hem.conv2d(x, 3, 64, name='1', use_batch_norm=False)
                e2 = hem.conv2d(e1, 64, 128, name='2')
                e3 = hem.conv2d(e2, 128, 256, name='3')
                e4 = hem.conv2d(e3, 256, 512, name='4')
                e5 = hem.conv2d(e4, 512, 512, name='5')
                e6 = hem.conv2d(e5, 512, 512, name='6')
                e7 = hem.conv2d(e6, 512, 512, name='7')
                e8 = hem.conv2d(e7, 512, 512, name='8')
        with arg_scope([hem.deconv2d, hem.conv2d], reuse=reuse, use_batch_norm=True, filter_size=4, stride=2,
                       init=init, activation=hem.fused_lrelu(0.0)):
            with variable_scope('decoder'):
                y = hem.deconv2d(e8, 512, 512, name='1', dropout=args.dropout)
                y = E.concat_channels([y, e7])
                y = hem.deconv2d(y, 1024, 512, name='2', dropout=args.dropout)
                y = E.concat_channels([y, e6])
                y = hem.deconv2d(y, 1024, 512, name='3', dropout=args.dropout)
                y = E.concat_channels([y, e5])
                y = hem.deconv2d(y, 1024, 512, name='4')
                y = E.concat_channels([y, e4])
                y = hem.deconv2d(y, 1024, 256, name='5')
                y = E.concat_channels([y, e3])
                y = hem.deconv2d(y, 512, 128, name='6')
                y = E.concat_channels([y, e2])
                y = hem.deconv2d(y, 256, 64, name='7')
                y = E.concat_channels([y, e1])
                y = hem.deconv2d(y, 128, 1, name='8', activation=tanh)
        return y

    @staticmethod
    def discriminator(x, y, args, reuse=False):
        """PatchGAN logits [B,8,8,1] (the sigmoid of pix2pix.py:259 only feeds summaries)."""
        init = lambda: random_normal_initializer(mean=0, stddev=0.02)
        with arg_scope([hem.conv2d], reuse=reuse, use_batch_norm=args.batch_norm_disc,
                       activation=hem.fused_lrelu(0.2), init=init, filter_size=4, stride=2):
            x_y = E.concat_channels([x, y])
            h = hem.conv2d(x_y, 4, 64, name='m1', use_batch_norm=False)
            h = hem.conv2d(h, 64, 128, name='m2')
            h = hem.conv2d(h, 128, 256, name='m3')
            h = hem.conv2d(h, 256, 512, name='m4')
            h = hem.conv2d(h, 512, 1, name='m5', activation=None)
        return h

    @staticmethod
    def loss(d_real_logits, d_fake_logits, g, x_depth, args):
        """pix2pix.py:263-304; returns the dict of the 'losses' collection."""
        g01 = hem.rescale(g, (-1, 1), (0, 1))
        y01 = hem.rescale(x_depth, (-1, 1), (0, 1))
        n = d_fake_logits.numel
        g_fake = E.eltloss(d_fake_logits, None, 4, label=1.0, scale=1.0 / n)
        l1 = E.eltloss(g01, y01, 0, scale=1.0 / g01.numel)
        g_total = g_fake
        if args.add_l1:
            g_total = E.add_scalars(g_fake, E.scale_scalar(l1, 10.0))
        d_real = E.eltloss(d_real_logits, None, 4, label=1.0, scale=1.0 / n)
        d_fake = E.eltloss(d_fake_logits, None, 4, label=0.0, scale=1.0 / n)
        d_total = E.add_scalars(d_real, d_fake)
        with E.recording(False):
            rmse = hem.rmse(y01, g01)
        return {'l1': l1, 'g_fake': g_fake, 'g_total': g_total, 'd_real': d_real, 'd_fake': d_fake,
                'd_total': d_total, 'rmse': rmse}
