"""Two eager IWGAN iterations at the bench config (for the ncu launch list: skip the first)."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200gan  # noqa
from b200gan import engine as E, session as S
from b200gan.models import gan as G
args = argparse.Namespace(model="iwgan", batch_size=512, latent_size=200, n_disc_train=5, optimizer="adam", lr=1e-4, beta1=0.5, beta2=0.9)
sess = S.Session(seed=0, noise_seed=1234)
sess.use_graphs = False
x = S.Input(512, (32, 32, 3), slots=6)
train = G.gan(x, args)
x.ring.copy_(torch.rand((6, 512, 32, 32, 3), device="cuda"))
for i in range(2):
    l0 = E.S.launches
    print(train(sess, args), "launches", E.S.launches - l0, flush=True)
