"""Debug aid: decompose the IWGAN critic-step gradient into (a) real/fake path, (b) first-order GP
gradient, (c) second-order GP term, and compare each with torch-CPU autograd over the oracle."""
import argparse
import sys
import os
from collections import OrderedDict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200gan  # noqa
from b200gan import engine as E, session as S, _capi as K
from b200gan.models import gan as G
from b200gan.ops.layers import variable_scope, flatten
from oracle import models as OM
from tests.parity import bf16_round, rel_err

H, C, L, B = 32, 3, int(os.environ.get("DBG_L", "16")), int(os.environ.get("DBG_B", "8"))
args = argparse.Namespace(model="iwgan", batch_size=B, latent_size=L, n_disc_train=1, optimizer="adam", lr=1e-4,
                          beta1=0.5, beta2=0.9)
sess = S.Session(seed=0)
sess.use_graphs = False
x_in = S.Input(B, (H, H, C), slots=1)
train = G.gan(x_in, args)
store = sess.store
gs, ds = OM.gan_param_specs("iwgan", H, C, L)
p = OM.init_params(OrderedDict(list(gs.items()) + list(ds.items())), 0)
for k_ in p:
    p[k_] = bf16_round(p[k_])
store.load(p)
gen = torch.Generator().manual_seed(1)
x01 = bf16_round(torch.rand(B, H, H, C, generator=gen))
z = bf16_round(torch.randn(B, L, generator=gen))
alpha = torch.rand(B, 1, generator=gen)
x_in.feed(0, x01.cuda())
d_params = store.collection("discriminator")
dnames = [n for n in store.params if n.startswith("discriminator")]


def report(tag, ref):
    for n in dnames:
        prm = store.params[n]
        got = prm.g32.reshape(prm.shape).float().cpu()
        want = ref[n]
        print("  %-14s %-36s err %.3e  |got| %.3e |want| %.3e" % (tag, n, rel_err(got, want) if want.norm() > 1e-9 else float((got - want).abs().max()), got.norm(), want.norm()))


def oracle_parts():
    q = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in p.items())
    x = 2 * (x01.reshape(B, -1) - 0.5)
    g = OM.generator(q, z, H, C, L).detach()
    d_real = OM.discriminator(q, x, H, C, L, "iwgan")
    d_fake = OM.discriminator(q, g, H, C, L, "iwgan")
    interp = (x + alpha * (g - x)).detach().requires_grad_(True)
    d_int = OM.discriminator(q, interp, H, C, L, "iwgan")
    grads = torch.autograd.grad(d_int.sum(), interp, create_graph=True)[0]
    ss = torch.sum(grads ** 2)
    gp = 10.0 * (torch.sqrt(ss) - 1) ** 2
    wd = d_fake.mean() - d_real.mean()
    dq = [q[n] for n in dnames]
    ga = torch.autograd.grad(wd, dq, retain_graph=True, allow_unused=True)
    gb = torch.autograd.grad(gp, dq, allow_unused=True)
    za = lambda gl: OrderedDict((n, torch.zeros_like(q[n]) if v is None else v) for n, v in zip(dnames, gl))
    return za(ga), za(gb), grads.detach(), float(ss), g


ref_a, ref_b, ref_grad, ref_ss, g_ref = oracle_parts()

# ---------------- (a) real/fake path only
sess.begin_step(); x_in.reset(); sess.noise_queue = [z.clone()]
for grp in store.groups:
    grp.zero_grad()
with E.recording(True, active=d_params):
    xr = flatten(E.affine(x_in.next(), 2.0, -1.0))
    with variable_scope("generator"), E.recording(False):
        g = G.generator(B, L, args, H, C)
    with variable_scope("discriminator"):
        d_real = G.discriminator(xr, args, H, C)
        d_fake = G.discriminator(g, args, H, C, reuse=True)
    gl, dl = E.wgan_losses(d_real, d_fake)
E.backward([(dl, None)])
torch.cuda.synchronize()
print("g (fake image) err", rel_err(g.torch().float(), g_ref))
report("real/fake", ref_a)

# ---------------- (b) first-order GP gradient and (c) second-order term
sess.begin_step(); sess.noise_queue = [alpha.clone()]
for grp in store.groups:
    grp.zero_grad()
with E.recording(True, active=d_params):
    with variable_scope("discriminator"):
        ss = G.gradient_penalty(xr, g, args, H, C, True)
    zero = E.Tensor(torch.zeros(B, device="cuda"));
    gl, dl = E.wgan_losses(zero, E.Tensor(torch.zeros(B, device="cuda")), ss, 10.0)
torch.cuda.synchronize()
print("sumsq got %.6e want %.6e" % (float(ss.buf.item()), ref_ss))
node = ss.node
gradT = node.inputs[0]
print("first-order grad err", rel_err(gradT.torch().float().reshape(B, -1), ref_grad))
E.backward([(dl, None)])
torch.cuda.synchronize()
report("gp-2nd-order", ref_b)
