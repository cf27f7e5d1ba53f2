"""Debug aid: print the loss trace of N iterations for a config (env: L, B, ITERS, GRAPHS)."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200gan  # noqa
from b200gan import session as S
from b200gan.models import gan as G
L = int(os.environ.get("L", 200)); B = int(os.environ.get("B", 512)); iters = int(os.environ.get("ITERS", 12))
args = argparse.Namespace(model="iwgan", batch_size=B, latent_size=L, n_disc_train=5, optimizer="adam", lr=1e-4, beta1=0.5, beta2=0.9)
sess = S.Session(seed=0, noise_seed=1234)
sess.use_graphs = os.environ.get("GRAPHS", "1") == "1"
x = S.Input(B, (32, 32, 3), slots=6)
train = G.gan(x, args)
gen = torch.Generator(device="cuda").manual_seed(1234)
REPEAT = int(os.environ.get("REPEAT", "0"))
pool = torch.rand((max(REPEAT, 1), 6, B, 32, 32, 3), generator=gen, device="cuda")
for i in range(iters):
    x.ring.copy_(pool[i % REPEAT] if REPEAT else torch.rand((6, B, 32, 32, 3), generator=gen, device="cuda"))
    print(L, B, i, train(sess, args), flush=True)
