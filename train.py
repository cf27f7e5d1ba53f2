#!/usr/bin/env python
"""train.py — the reference's CLI for the hot path (/root/reference/train.py:44-184 flags, 240-246 model
dispatch, 295-329 epoch loop), driving the B200 engine.

  python train.py --model iwgan --batch_size 512 --optimizer adam --lr 1e-4 --beta1 0.5 --beta2 0.9 \
                  --epochs 1 --epoch_size 5120 [--config FILE] [--n_gpus N via torchrun]

Same flags, `--config` files (whitespace `key value` tokens, CLI wins, train.py:25-37) and
`train_func(sess, args) -> {loss: float}` contract.  Differences, all outside the hot path: the input
pipeline is synthetic ([0,1] float32 batches of `--image_size`/`--channels`; TFRecord decoding is out of
scope), `--n_gpus` is realised as one process per GPU (launch with torchrun; each rank
is one tower with its own batch of `--batch_size`), summaries/TensorBoard are not produced.  Checkpoints
(3dgan_b200/checkpoint.py) are torch files keyed by the reference's TF variable / slot names; like the
reference's Supervisor the run restores the newest checkpoint found in `--dir`, `--epochs +n` trains n more
epochs, a checkpoint is written before the first step and after every epoch.
"""
import argparse
import os
import sys
import time
import uuid

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


class load_args_from_file(argparse.Action):
    """train.py:25-37 — `key value` tokens, `--` prepended when missing; explicit CLI flags win."""

    def __call__(self, parser, namespace, values, option_string=None):
        contents = values.read().split()
        for i in range(len(contents) // 2):
            if contents[i * 2][0:2] != '--':
                contents[i * 2] = '--' + contents[i * 2]
        data = parser.parse_args(contents, namespace=namespace)
        for k, v in vars(data).items():
            if v and k != option_string.strip('-'):
                setattr(namespace, k, v)


def build_parser():
    p = argparse.ArgumentParser(description="3dgan training step on B200")
    add = p.add_argument_group('Miscellaneous').add_argument
    add('--config', type=open, action=load_args_from_file)
    add('--seed', type=int, default=None)
    add('--n_gpus', type=int, default=1)
    add('--profile', default=False, action='store_true')
    add = p.add_argument_group('Training').add_argument
    add('--epochs', default='3')
    add('--batch_size', type=int, default=256)
    add('--epoch_size', type=int, default=-1)
    add('--examples', type=int, default=64)
    add('--dir', type=str, default='workspace/{}'.format(uuid.uuid4()))
    add('--n_disc_train', type=int, default=5)
    add = p.add_argument_group('Optimizer').add_argument
    add('--optimizer', type=lambda s: s.lower(), default='rmsprop')
    add('--lr', type=float, default=0.001)
    add('--loss', type=lambda s: s.lower(), default='l1')
    add('--momentum', type=float, default=0.01)
    add('--decay', type=float, default=0.9)
    add('--centered', default=False, action='store_true')
    add('--beta1', type=float, default=0.9)
    add('--beta2', type=float, default=0.999)
    add = p.add_argument_group('Model').add_argument
    add('--model', type=lambda s: s.lower(), default='iwgan')
    add('--latent_size', type=int, default=200)
    add = p.add_argument_group('Data').add_argument
    add('--dataset', type=lambda s: s.lower(), default='synthetic')
    add('--image_size', type=int, default=64)
    add('--channels', type=int, default=3)
    add('--resize', type=int, nargs=2, default=None)
    add('--shuffle', default=True)
    add('--buffer_size', type=int, default=10000)
    add('--grayscale', default=False, action='store_true')
    add('--cache_dir', default=None)
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    import torch
    import b200gan  # noqa: F401
    from b200gan import checkpoint as CK
    from b200gan import session as S
    from b200gan.models import MODEL_FUNCS

    if args.model not in MODEL_FUNCS:
        raise SystemExit("model '%s' is not on the accelerated path (%s)" % (args.model, sorted(MODEL_FUNCS)))
    sess = S.Session(seed=0 if args.seed is None else args.seed,
                     noise_seed=1234 if args.seed is None else args.seed)
    if sess.world > 1:
        sess.init_distributed("nccl")
    # save options to disk for later reference (train.py:200-208)
    if sess.rank == 0:
        os.makedirs(args.dir, exist_ok=True)
        with open(os.path.join(args.dir, 'options.config'), 'w') as f:
            for k in vars(args):
                f.write('{} {}\n'.format(k, getattr(args, k)))
    runs = MODEL_FUNCS[args.model][1](args)
    x = S.Input(args.batch_size, (args.image_size, args.image_size, args.channels), slots=runs)
    train_func = MODEL_FUNCS[args.model][0](x, args)                 # train.py:246
    n_examples = args.epoch_size if args.epoch_size > 0 else 100 * args.batch_size
    iter_per_epoch = max(1, n_examples // (args.batch_size * max(sess.world, 1)))   # train.py:221-224
    gen = torch.Generator(device="cuda").manual_seed(1234 + sess.rank)

    # Supervisor.managed_session: restore the newest checkpoint of --dir if there is one (train.py:254-259,276-278)
    current_step, current_epoch = 0, 0
    restored = CK.restore_latest(sess, args.dir)
    if restored is not None:
        path, current_step, current_epoch, extra = restored
        if extra.get("data_rng") is not None and sess.rank == 0:
            gen.set_state(extra["data_rng"])
        if sess.rank == 0:
            print("Restored %s (global_step %d, global_epoch %d)" % (path, current_step, current_epoch), flush=True)
    # `--epochs +n`: n more epochs on top of the restored ones (train.py:279-282)
    max_epochs = current_epoch + int(args.epochs[1:]) if args.epochs[0] == '+' else int(args.epochs)

    def checkpoint(tag, epoch):
        CK.save(sess, args.dir, tag, global_epoch=epoch, extra={"data_rng": gen.get_state()})

    if current_step == 0:                                             # train.py:288-292: parameters before any training
        checkpoint(0, 0)
    start = time.time()
    status = None
    for epoch in range(current_epoch, max_epochs):
        t0 = time.time()
        for i in range(iter_per_epoch):
            x.ring.copy_(torch.rand(x.ring.shape, generator=gen, device="cuda"))
            status = train_func(sess, args)                           # train.py:307
        current_epoch = epoch + 1                                     # increment_global_epoch, train.py:323-325
        if sess.rank == 0:
            dt = time.time() - t0
            print("Epoch %3d  %d it  %.2f batch/s  %s" % (current_epoch, iter_per_epoch, iter_per_epoch / dt,
                                                         {k: round(v, 5) for k, v in status.items()}), flush=True)
        checkpoint(current_epoch, current_epoch)                      # saver.save(..., global_step=global_epoch), train.py:329
    if sess.rank == 0:
        print("Training complete! Elapsed time: %ds" % int(time.time() - start))
    return status


if __name__ == '__main__':
    main()
