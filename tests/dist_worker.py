"""Worker of tests/test_multi_gpu.py: one rank of a data-parallel IWGAN run (launched with torch.distributed.run).

Every rank is one tower (util.py:54-77) with its own batch and noise; rank 0 also evaluates the CPU oracle for ALL
towers (their inputs are seeded by rank, so it can regenerate them) and checks
  1. the exchanged critic / generator gradients (flat bucket after the NCCL all-reduce x 1/n) against the oracle's
     per-variable tower mean, `average_gradients` (util.py:118-147), on the decision-injected fp32 oracle;
  2. the parameters after `iters` full iterations (n_disc critic updates + 1 generator update, Adam) against the
     oracle's multi-tower trajectory.
Writes a JSON report to --out (rank 0)."""
import argparse
import json
import os
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import b200gan  # noqa: E402,F401
from b200gan import engine as E  # noqa: E402
from b200gan import session as S  # noqa: E402
from b200gan.models import gan as gan_model  # noqa: E402
from oracle import models as OM  # noqa: E402
from oracle import tf_ops as OT  # noqa: E402
from tests import parity as P  # noqa: E402

H, C, L, B, N_DISC = 32, 3, 16, 16, 2


def tower_inputs(rank, run):
    g = torch.Generator().manual_seed(1000 * rank + run + 1)
    return (P.bf16_round(torch.rand(B, H, H, C, generator=g)), P.bf16_round(torch.randn(B, L, generator=g)),
            torch.rand(B, 1, generator=g))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--iters", type=int, default=2)
    a = ap.parse_args()
    sess = S.Session(seed=0)
    sess.use_graphs = False
    sess.init_distributed("nccl")
    world, rank = sess.world, sess.rank
    args = argparse.Namespace(model="iwgan", batch_size=B, latent_size=L, n_disc_train=N_DISC, optimizer="adam",
                              lr=1e-4, beta1=0.5, beta2=0.9)
    runs = N_DISC + 1
    x_in = S.Input(B, (H, H, C), slots=runs)
    train = gan_model.gan(x_in, args)
    store = sess.store
    gs, ds = OM.gan_param_specs("iwgan", H, C, L)
    p = OM.init_params(OrderedDict(list(gs.items()) + list(ds.items())), 0)
    for k in p:
        p[k] = P.bf16_round(p[k])
    store.load(p)
    report = {"world": world, "nccl_abi": sess.comm is not None, "overlap": sess.overlap_updates, "ok": True}

    # ---- 1. exchanged gradients of one critic run and one generator run
    d_group = [g for g in store.groups if g.name == "discriminator"][0]
    g_group = [g for g in store.groups if g.name == "generator"][0]
    traces = {}
    got = {}
    for mode, grp in (("d", d_group), ("g", g_group)):
        x01, z, alpha = tower_inputs(rank, 0)
        sess.begin_step()
        x_in.reset()
        x_in.feed(0, x01.cuda())
        sess.noise_queue = [z.clone(), alpha.clone()]
        for g_ in store.groups:
            g_.zero_grad()
        E.S.decisions = []
        gl, dl = train.tower(x_in.next(), mode)
        ex = sess.exchange(grp)
        ex.begin()
        E.backward([(dl if mode == "d" else gl, None)], on_ready=ex.on_ready)
        ex.finish()
        scale = ex.join()
        torch.cuda.synchronize()
        traces[mode] = P.gpu_decisions(E.S.decisions)
        E.S.decisions = None
        got[mode] = {n: (prm.logical(prm.g32).float() * scale).cpu().clone() for n, prm in store.params.items()
                     if prm.group is grp}
        report["buckets_" + mode] = len(ex.buckets)
    # rank 0 needs every tower's decisions: gather them through torch.distributed (control plane)
    for mode in ("d", "g"):
        flat = torch.cat([m.reshape(-1) for _, m in traces[mode]]).to(torch.int8).cuda()
        parts = [torch.empty_like(flat) for _ in range(world)]
        torch.distributed.all_gather(parts, flat)
        if rank == 0:
            shapes = [(k, m.shape) for k, m in traces[mode]]
            worst = 0.0
            mean = None
            for r in range(world):
                dec, o = [], 0
                fl = parts[r].cpu().float()
                for k, shp in shapes:
                    n = int(torch.tensor(shp).prod())
                    dec.append((k, fl[o:o + n].reshape(shp)))
                    o += n
                x01, z, alpha = tower_inputs(r, 0)
                with OT.inject_decisions(dec) as inj:
                    ref = OM.gan_grads(p, x01, z, alpha, "iwgan", H, C, L)
                assert not inj.queue
                mean = ref["grads"] if mean is None else OrderedDict((k, mean[k] + v) for k, v in ref["grads"].items())
            for n, g_ in got[mode].items():
                want = mean[n] / world                         # average_gradients: per-variable tower mean
                wn = float(want.norm())
                if wn < 1e-6:
                    continue
                e = float((g_ - want).norm()) / wn
                worst = max(worst, e)
                # small towers: generator variables sit below small-batch batch norms (tests/parity.py storage_noise_floor)
                tol = 8e-2 if n.startswith("generator/") else 3e-2
                if e > tol:
                    report["ok"] = False
                    report.setdefault("bad", []).append((mode, n, e))
            report["worst_grad_err_" + mode] = worst

    # ---- 2. trajectory: `iters` iterations on every tower, parameters vs the oracle's multi-tower Adam
    store.load(p)
    for g_ in store.groups:
        g_.zero_grad(); g_.m.zero_(); g_.v.zero_(); g_.step.zero_()
    ref_p = OrderedDict((k, v.clone()) for k, v in p.items())
    g_opt = OM.AdamState(ref_p, list(gs), 1e-4, 0.5, 0.9)
    d_opt = OM.AdamState(ref_p, list(ds), 1e-4, 0.5, 0.9)
    for it in range(a.iters):
        ins = [tower_inputs(rank, 10 + it * runs + k) for k in range(runs)]
        for k, (x01, _, _) in enumerate(ins):
            x_in.feed(k, x01.cuda())
        q = []
        for _, z, alpha in ins:
            q += [z.clone(), alpha.clone()]
        sess.noise_queue = q
        sess.begin_step()
        out = train.iteration()
        torch.cuda.synchronize()
        if rank == 0:
            with OT.store_bf16(True):
                for k in range(runs):
                    tg = []
                    for r in range(world):
                        x01, z, alpha = tower_inputs(r, 10 + it * runs + k)
                        tg.append(OM.gan_grads(ref_p, x01, z, alpha, "iwgan", H, C, L)["grads"])
                    avg = OrderedDict((n, sum(t[n] for t in tg) / world) for n in tg[0])
                    (d_opt if k < N_DISC else g_opt).apply(ref_p, avg)
    if rank == 0:
        worst = 0.0
        for n, prm in store.params.items():
            d_ref = ref_p[n] - p[n]
            if float(d_ref.norm()) < 1e-9 or (n.startswith("generator/vars/") and n.endswith("/bias") and "dc3" not in n):
                continue                                      # biases under batch norm: Adam random-walks them (App. C #6)
            d_got = prm.logical(prm.p32).float().cpu() - p[n]
            worst = max(worst, float((d_got - d_ref).norm() / d_ref.norm()))
        report["worst_displacement_err"] = worst
        report["losses"] = {k: float(v.item()) for k, v in out.items()}
        if worst > 0.25:
            report["ok"] = False
    # every rank must hold identical parameters after the exchange
    for g_ in store.groups:
        ref_t = g_.p32.clone()
        torch.distributed.broadcast(ref_t, 0)
        same = bool(torch.equal(ref_t, g_.p32))
        flag = torch.tensor([int(same)], device="cuda")
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
        if rank == 0 and int(flag.item()) != 1:
            report["ok"] = False
            report["ranks_diverged"] = g_.name
    if rank == 0:
        json.dump(report, open(a.out, "w"))
        print(json.dumps(report))
    torch.cuda.synchronize()
    torch.distributed.barrier()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
