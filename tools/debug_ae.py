"""Debug aid: per-layer activation and activation-gradient errors of the cnn autoencoder vs the oracle."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200gan  # noqa
from b200gan import engine as E, session as S
from b200gan.models import MODEL_FUNCS
from oracle import models as OM, tf_ops as OT
from tests.parity import bf16_round, rel_err

H, C, L, B = 28, 1, int(os.environ.get("DBG_L", 200)), int(os.environ.get("DBG_B", 64))
EMU = os.environ.get("EMU", "1") == "1"
args = argparse.Namespace(model="cnn", batch_size=B, latent_size=L, n_disc_train=1, optimizer="adam", lr=1e-3, beta1=0.9, beta2=0.999)
sess = S.Session(seed=0); sess.use_graphs = False
x_in = S.Input(B, (H, H, C), slots=1)
train = MODEL_FUNCS["cnn"][0](x_in, args)
specs, sizes = OM.ae_param_specs("cnn", H, C, L)
p = OM.init_params(specs, 0)
for k in p: p[k] = bf16_round(p[k])
sess.store.load(p)
gen = torch.Generator().manual_seed(3)
x01 = bf16_round(torch.rand(B, H, H, C, generator=gen))
x_in.feed(0, x01.cuda())

# ---- oracle with retained activations
acts = []
def keep(t):
    t.retain_grad(); acts.append(t); return t
with OT.store_bf16(EMU):
    q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    x = OT.stored(2 * (x01 - 0.5))
    h = x
    for i, s in enumerate([2, 2, 2, 2, 1, 1], 1):
        h = keep(OT.conv2d(h, q["encoder/vars/c%d/weights" % i], q["encoder/vars/c%d/bias" % i], s, None, "lrelu"))
    z = keep(OT.dense(h.reshape(B, -1), q["latent/vars/d1/weights"], q["latent/vars/d1/bias"]))
    s_ = sizes[-1]
    h = keep(OT.dense(z, q["decoder/vars/d1/weights"], q["decoder/vars/d1/bias"], None, "relu"))
    h = h.reshape(-1, s_, s_, 32)
    h = keep(OT.conv2d(h, q["decoder/vars/c1/weights"], q["decoder/vars/c1/bias"], 1, None, "relu"))
    h = keep(OT.conv2d(h, q["decoder/vars/c2/weights"], q["decoder/vars/c2/bias"], 1, None, "relu"))
    for i in range(1, 5):
        o = sizes[4 - i]
        h = keep(OT.deconv2d(h, q["decoder/vars/dc%d/weights" % i], q["decoder/vars/dc%d/bias" % i], 2, None, "tanh" if i == 4 else "relu", out_hw=(o, o)))
    loss = torch.mean(torch.abs(x - h))
    loss.backward()

# ---- ours, recording every node's delivered output gradient
sess.begin_step(); x_in.reset(); sess.store.groups[0].zero_grad()
ours_acts = []
orig_conv = E.conv_like
def conv_spy(direction, x, W, geom, **kw):
    out = orig_conv(direction, x, W, geom, **kw)
    if E.S.recording and kw.get("bias") is not None:
        ours_acts.append(out)
    return out
E.conv_like = conv_spy
import b200gan.ops.layers as LY
out = train.tower(x_in.next())
E.conv_like = orig_conv
delivered = {}
orig_bw = E.backward
# instrument: wrap each node.bw to capture gouts
for t in ours_acts:
    node = t.node
    def mk(node, t):
        old = node.bw
        def bw(gouts):
            delivered[id(t)] = gouts[0]
            return old(gouts)
        node.bw = bw
    mk(node, t)
E.backward([(out, None)])
torch.cuda.synchronize()
print("loss ours %.7f oracle %.7f" % (float(out.buf.item()), float(loss)))
names = ["enc c1", "enc c2", "enc c3", "enc c4", "enc c5", "enc c6", "latent d1", "dec d1", "dec c1", "dec c2", "dec dc1", "dec dc2", "dec dc3", "dec dc4"]
assert len(ours_acts) == len(acts), (len(ours_acts), len(acts))
for nme, ta, tb in zip(names, ours_acts, acts):
    a_err = rel_err(ta.torch().float().reshape(tb.shape), tb)
    g = delivered.get(id(ta))
    # ours delivers the gradient already multiplied by act'(a); oracle's retained grad is w.r.t. the activation
    if "dc4" in nme: d = 1 - tb.detach() ** 2
    elif nme.startswith("enc"): d = torch.where(tb.detach() > 0, torch.ones_like(tb), torch.full_like(tb, 0.2))
    elif nme == "latent d1": d = torch.ones_like(tb)
    else: d = (tb.detach() > 0).float()
    want = tb.grad * d
    g_err = rel_err(g.torch().float().reshape(tb.shape), want) if g is not None else float("nan")
    print("%-10s act err %.3e   grad err %.3e   |grad| %.3e  zeros ours %.3f oracle %.3f" % (nme, a_err, g_err, float(want.norm()),
          float((ta.torch().float() == 0).float().mean()), float((tb == 0).float().mean())))
