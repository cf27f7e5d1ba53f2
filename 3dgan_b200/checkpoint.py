"""Checkpoints of the training state, keyed by the reference's TensorFlow variable names.

What `tf.train.Supervisor(logdir=args.dir, saver=tf.train.Saver(max_to_keep=0))` kept for the reference
(train.py:254-259 auto-restore, 279-282 `--epochs +n`, 288-292 the step-0 checkpoint, 329 one checkpoint per
epoch tagged with global_epoch):

  variables        `generator/vars/dc1/weights`, `discriminator/vars/c2/bias`, `.../BatchNorm_1/beta`, ...
  optimizer slots  `<variable>/Adam`, `<variable>/Adam_1` (m, v); `<variable>/RMSProp`, `/RMSProp_1`; `/Momentum`;
                   `/Adagrad`; `/Adadelta`, `/Adadelta_1`; `/Ftrl`, `/Ftrl_1`  (TF-1.x slot naming)
  per optimizer    its step count t (TF stores beta1^t / beta2^t as `beta1_power` / `beta2_power`: both are written)
  batch norm       `<scope>/BatchNorm[_k]/moving_mean`, `/moving_variance`  (non-trainable, models/gan.py:69-70)
  counters         `global_step` (one increment per apply_gradients), `global_epoch`
  + what a bit-faithful resume of THIS implementation needs: the Philox draw counter and the data generator state.

Files are torch-serialised dicts `checkpoint-<tag>.pt` plus a one-line index file `checkpoint` naming the newest
(TF's `checkpoint` text file plays the same role).  Values are stored at their LOGICAL (TF) shapes: the zero
channel padding of the physical buffers is stripped on save and restored on load.
"""
import os
from collections import OrderedDict

import torch

from . import _capi as K
from . import engine as E

FORMAT = "b200gan-checkpoint-v1"
INDEX = "checkpoint"


def _pad(t, shape):
    if tuple(t.shape) == tuple(shape):
        return t
    out = torch.zeros(shape, dtype=t.dtype, device=t.device)
    out[tuple(slice(0, d) for d in t.shape)] = t
    return out


def state_dict(sess, global_epoch=0, extra=None):
    store = sess.store
    ck = OrderedDict(format=FORMAT)
    ck["variables"] = store.state_dict()
    slots, optim = OrderedDict(), OrderedDict()
    gstep = 0
    for gi, g in enumerate(store.groups):
        t = int(g.step.item())
        gstep += t
        suffix = "" if gi == 0 else "_%d" % gi
        optim[g.name] = {"step": t, "optimizer": g.cfg["optimizer"]}
        if g.kind == K.OPT_ADAM:                          # TF keeps the bias-correction powers, not t
            optim[g.name]["optimizers/beta1_power" + suffix] = float(g.cfg["beta1"]) ** (t + 1)
            optim[g.name]["optimizers/beta2_power" + suffix] = float(g.cfg["beta2"]) ** (t + 1)
        for slot, buf in g.slots():
            for p in g.params:
                slots["%s/%s" % (p.name, slot)] = p.logical(buf[p.offset:p.offset + p.numel].detach()).cpu().clone()
    ck["slots"], ck["optimizers"] = slots, optim
    ck["state"] = OrderedDict((n, sv.logical(sv.buf.detach()).cpu().clone()) for n, sv in store.state.items())
    ck["global_step"], ck["global_epoch"] = gstep, int(global_epoch)
    ck["noise_counter"] = None if sess.counter is None else int(sess.counter.item())
    ck["extra"] = dict(extra or {})
    return ck


def load_state_dict(sess, ck):
    """Restore everything state_dict() saved into an already-built (finalized) session."""
    if ck.get("format") != FORMAT:
        raise K.B200Error("not a %s checkpoint" % FORMAT)
    store = sess.store
    missing = set(store.params) ^ set(ck["variables"])
    if missing:
        raise K.B200Error("checkpoint and model disagree on variables: %s" % sorted(missing)[:4])
    store.load(ck["variables"])
    for g in store.groups:
        meta = ck["optimizers"].get(g.name)
        if meta is None or meta["optimizer"] != g.cfg["optimizer"]:
            raise K.B200Error("checkpoint was written by optimizer %s, this run uses %s for group %s"
                              % (meta and meta["optimizer"], g.cfg["optimizer"], g.name))
        g.step.fill_(int(meta["step"]))
        for slot, buf in g.slots():
            for p in g.params:
                v = ck["slots"]["%s/%s" % (p.name, slot)]
                buf[p.offset:p.offset + p.numel].copy_(_pad(v.reshape(p.logical_shape).to(torch.float32), p.shape).reshape(-1))
        g.g32.zero_()
    for n, sv in store.state.items():
        sv.buf.copy_(_pad(ck["state"][n].reshape(sv.logical_shape).to(torch.float32), sv.shape).reshape(-1))
    if ck.get("noise_counter") is not None:
        if sess.counter is None:
            sess.counter = torch.zeros(1, dtype=torch.int64, device=sess.device)
        sess.counter.fill_(int(ck["noise_counter"]))
    return ck["global_step"], ck["global_epoch"], ck.get("extra", {})


def save(sess, directory, tag, global_epoch=0, extra=None):
    """Write `<directory>/checkpoint-<tag>.pt` and point the index file at it.  With several ranks: batch norm moving
    averages are taken from the LAST rank (the reference runs the last tower's UPDATE_OPS, models/gan.py:69-70);
    rank 0 writes."""
    if sess.world > 1 and sess.store.state_buf is not None and sess.store.state:
        if sess.cuda:
            E.begin(sess.device)
        sess.broadcast(sess.store.state_buf, sess.world - 1)
    if sess.rank != 0:
        return None
    os.makedirs(directory, exist_ok=True)
    name = "checkpoint-%s.pt" % tag
    tmp = os.path.join(directory, name + ".tmp")
    torch.save(state_dict(sess, global_epoch, extra), tmp)
    os.replace(tmp, os.path.join(directory, name))
    with open(os.path.join(directory, INDEX), "w") as f:
        f.write(name + "\n")
    return os.path.join(directory, name)


def latest(directory):
    """Path of the newest checkpoint of `directory` (what Supervisor.managed_session would restore), or None."""
    idx = os.path.join(directory, INDEX)
    if not os.path.exists(idx):
        return None
    name = open(idx).read().strip()
    path = os.path.join(directory, name)
    return path if name and os.path.exists(path) else None


def restore_latest(sess, directory):
    path = latest(directory)
    if path is None:
        return None
    ck = torch.load(path, map_location="cpu", weights_only=False)
    return (path,) + tuple(load_state_dict(sess, ck))
