"""CPU: host-side logic of the hot path — graph-construction (dry) pass, variable naming, kernel
schedule of the critic step, arg_scope, tower slicing and the data-parallel exchange (gloo, world 2)."""
import argparse
import os
from collections import OrderedDict

import pytest
import torch

import b200gan  # noqa: F401
from b200gan import engine as E
from b200gan import session as S
from b200gan.models import gan as gan_model
from b200gan.ops import layers as L
from b200gan.ops.activations import lrelu
from b200gan.ops.arg_scope import arg_scope
from oracle import models as OM
from oracle import tf_ops as OT


def _args(model="iwgan", B=8, Lz=16):
    return argparse.Namespace(model=model, batch_size=B, latent_size=Lz, n_disc_train=5, optimizer="adam", lr=1e-4,
                              beta1=0.5, beta2=0.9)


@pytest.mark.parametrize("model,H", [("iwgan", 32), ("iwgan", 64), ("gan", 32), ("wgan", 64)])
def test_build_pass_creates_reference_variables(model, H):
    """The dry graph-construction pass creates exactly the TF variables (names, shapes, order) the
    reference would (oracle.gan_param_specs restates models/gan.py + ops/layers.py naming)."""
    sess = S.Session()
    x = S.Input(8, (H, H, 3), slots=6)
    gan_model.gan(x, _args(model))
    gs, ds = OM.gan_param_specs(model, H, 3, 16)
    want = OrderedDict(list(gs.items()) + list(ds.items()))
    got = OrderedDict((n, p.logical_shape) for n, p in sess.store.params.items())
    assert set(got) == set(want)
    for n in want:
        assert tuple(want[n]) == got[n], n
    # creation order inside each scope follows the reference's layer order
    assert [n for n in got if n.startswith("generator")][:3] == list(gs)[:3]


def _trace_launches(fn):
    calls = []
    orig = E.launch

    def rec(name, *a, **k):
        calls.append(name)
        return 0
    E.launch = rec
    try:
        fn()
    finally:
        E.launch = orig
    return calls


def test_critic_step_kernel_schedule():
    """One IWGAN critic run = 3 critic forwards + first-order GP backward + second-order sweep + real/fake
    backward; count the conv-family launches of each kind (~10 D_fwd-equivalents)."""
    sess = S.Session()
    x = S.Input(8, (32, 32, 3), slots=6)
    train = gan_model.gan(x, _args())
    E.S.dry = True
    try:
        sess.store.begin_pass()

        def run():
            gl, dl = train.tower(x.next(), "d")
            E.backward([(dl, None)])
        calls = _trace_launches(run)
    finally:
        E.S.dry = False
    n_f = calls.count("b200_conv2d_fprop")
    n_d = calls.count("b200_conv2d_dgrad")
    # (image-side c1: the real / fake filter gradients also yield the bias gradient -- b200_conv2d_wgrad_bias)
    n_w = calls.count("b200_conv2d_wgrad") + calls.count("b200_conv2d_wgrad_bias")
    assert calls.count("b200_conv2d_wgrad_bias") == 2, calls
    # fprop: G (fc1) + 3 critic passes x 3 convs + second-order 3  = 1 + 9 + 3
    assert n_f == 13, calls
    # dgrad: G's 3 deconvs + first-order GP chain (3) + real and fake paths (c3, c2 each)
    assert n_d == 3 + 3 + 4, calls
    # wgrad: 3 convs x (real, fake, second-order)
    assert n_w == 9, calls
    assert calls.count("b200_gemv_rows") == 3 and calls.count("b200_wgan_loss") == 1


def test_generator_step_does_not_touch_critic_gradients():
    sess = S.Session()
    x = S.Input(8, (32, 32, 3), slots=6)
    train = gan_model.gan(x, _args())
    E.S.dry = True
    try:
        sess.store.begin_pass()
        touched = []
        orig = E.launch

        def rec(name, *a, **k):
            touched.append(name)
            return 0
        E.launch = rec
        gl, dl = train.tower(x.next(), "g")
        E.backward([(gl, None)])
        E.launch = orig
    finally:
        E.S.dry = False
    # generator run: wgrad only for fc1 + 3 deconvs; the GP path is evaluated (first order) but not differentiated
    assert touched.count("b200_conv2d_wgrad") == 4
    assert touched.count("b200_bn_bwd") == 3


def test_arg_scope_overrides_defaults_and_nests():
    seen = {}

    @L.add_arg_scope
    def layer(x, a=1, b=2):
        seen["v"] = (a, b)
    with arg_scope([layer], a=10):
        layer(0)
        assert seen["v"] == (10, 2)
        with arg_scope([layer], b=20):
            layer(0, a=11)
            assert seen["v"] == (11, 20)
        layer(0)
        assert seen["v"] == (10, 2)
    layer(0)
    assert seen["v"] == (1, 2)


def test_activation_tags_and_same_padding():
    assert lrelu.b200_act[0] == b200gan._capi.ACT_LRELU and abs(lrelu.b200_act[1] - 0.2) < 1e-9
    for size, k, s in [(64, 5, 2), (28, 5, 2), (7, 5, 2), (256, 4, 2), (8, 1, 1)]:
        out, before, _ = OT.same_pad(size, k, s)
        assert E.same_pad(size, k, s) == (out, before)


def test_no_gpu_means_loud_failure():
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    sess = S.Session()
    x = S.Input(8, (32, 32, 3), slots=6)
    train = gan_model.gan(x, _args())
    with pytest.raises(RuntimeError):
        train(sess, _args())


# --------------------------------------------------------------------------------------------- gloo, world 2
def _dp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    sess = S.Session()
    sess.cuda = False
    sess.init_distributed("gloo")

    class G:       # stands in for variables.Group: the exchange only needs the flat gradient bucket
        g32 = torch.arange(8, dtype=torch.float32) * (rank + 1)
    scale = sess.all_reduce_grads(G)
    avg = G.g32 * scale
    # Session.Exchange: buckets cut in reverse creation order, each sent once all of its variables are final
    class Prm:
        def __init__(self, off, n):
            self.offset, self.numel = off, n

    class Grp:
        params = [Prm(0, 64), Prm(64, 192), Prm(256, 64), Prm(320, 704)]
        size = 1024
        g32 = torch.arange(1024, dtype=torch.float32) * (rank + 1)
    ex = S.Exchange(sess, Grp, bucket_bytes=1024)             # >= 256 elements per bucket
    bounds = [(lo, hi) for lo, hi, _ in ex.buckets]
    ex.begin()
    sent = []
    for prm in reversed(Grp.params):
        ex.on_ready(prm)
        sent.append(ex.next)
    ex.finish()
    ex_scale = ex.join()
    ex_ok = bool(torch.equal(Grp.g32 * ex_scale, torch.arange(1024, dtype=torch.float32) * 1.5))
    # ops.input.batch_slice: tower r takes rows [r*B, (r+1)*B)
    from b200gan.ops.input import batch_slice
    glob = E.Tensor(torch.arange(4 * world, dtype=torch.float32).reshape(4 * world, 1))
    sl = batch_slice(glob, 4, rank).torch().flatten().tolist()
    if rank == 0:
        out.put((avg.tolist(), scale, sl, bounds, sent, ex_ok))
    else:
        out.put((None, scale, sl, bounds, sent, ex_ok))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_exchange_is_tower_mean_gloo_world2():
    """average_gradients (util.py:118-147) == sum all-reduce x 1/n; batch_slice == tower rows."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    avg = [r[0] for r in res if r[0] is not None][0]
    towers = [[torch.arange(8, dtype=torch.float32) * (r + 1)] for r in range(2)]
    want = OT.average_gradients(towers)[0].tolist()
    assert avg == pytest.approx(want)
    assert all(abs(r[1] - 0.5) < 1e-9 for r in res)
    slices = sorted(r[2] for r in res)
    assert slices == [[0.0, 1.0, 2.0, 3.0], [4.0, 5.0, 6.0, 7.0]]
    for r in res:
        assert r[3] == [(320, 1024), (64, 320), (0, 64)]      # last variables first; the small tail is its own bucket
        assert r[4] == [1, 1, 2, 3]                            # a bucket leaves when its LAST variable is final
        assert r[5]                                            # every element exchanged exactly once: tower mean


def test_pix2pix_build_pass_matches_reference_variables():
    """54.4 M generator + 2.77 M discriminator parameters, TF names incl. the 'enocder' typo (pix2pix.py:182)."""
    from b200gan.models import pix2pix
    from oracle import pix2pix as OP
    a = argparse.Namespace(batch_size=2, n_disc_train=1, optimizer="adam", lr=1e-4, beta1=0.5, beta2=0.999,
                           batch_norm_gen=False, batch_norm_disc=False, add_l1=True, dropout=0, noise=[])
    sess = S.Session()
    pix2pix((S.Input(2, (256, 256, 3), slots=3), S.Input(2, (256, 256, 1), slots=3)), a)
    gs, ds = OP.param_specs()
    want = OrderedDict(list(gs.items()) + list(ds.items()))
    got = OrderedDict((n, p.logical_shape) for n, p in sess.store.params.items())
    assert list(got) == list(want)
    assert all(tuple(want[n]) == got[n] for n in want)
    assert abs(sum(p.numel for n, p in sess.store.params.items() if n.startswith("generator")) - 54.408e6) < 1e3


def test_small_output_conv_routes_to_simt_backward():
    from b200gan import _capi
    g = _capi.ConvGeom(N=2, H=16, W=16, Cin=512, Ho=8, Wo=8, Cout=1, k=4, stride=2, pad_t=1, pad_l=1)
    assert _capi.route(g, 0) == 1 and _capi.route(g, 1) == 3 and _capi.route(g, 2) == 3


def test_train_cli_flags_and_config_file(tmp_path):
    """train.py keeps the reference's flags and `--config` file semantics (train.py:25-37: whitespace
    `key value` tokens, `--` prepended, explicit CLI flags win)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("b200_train", os.path.join(os.path.dirname(os.path.dirname(
        os.path.abspath(__file__))), "train.py"))
    tr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tr)
    cfg = tmp_path / "iwgan.config"
    cfg.write_text("model\t\tiwgan\nepochs\t\t20\nbatch_size \t256\nn_gpus \t\t2\noptimizer\tadam\nlr\t\t1e-4\nbeta1\t\t0.5\nbeta2\t\t0.9\n")
    a = tr.build_parser().parse_args(["--config", str(cfg)])
    assert (a.model, a.epochs, a.batch_size, a.n_gpus, a.optimizer, a.lr, a.beta1, a.beta2) == \
        ("iwgan", "20", 256, 2, "adam", 1e-4, 0.5, 0.9)
    d = tr.build_parser().parse_args([])
    assert (d.optimizer, d.lr, d.momentum, d.decay, d.latent_size, d.n_disc_train, d.batch_size) == \
        ("rmsprop", 0.001, 0.01, 0.9, 200, 5, 256)          # reference defaults, train.py:87-153


def test_channel_padding_is_layout_only():
    """200-channel layers are stored with 8 zero channels: the TF variables keep the reference's
    names and shapes (logical view), only the physical arrays grow; initialisation / load / state_dict go
    through the logical block and leave the padding at exactly zero."""
    import torch
    from b200gan.ops import layers as L
    from b200gan import variables as V
    assert [L.physical_channels(c) for c in (3, 8, 24, 72, 100, 200, 208, 400, 800)] == [3, 8, 32, 80, 112, 208, 208, 400, 800]
    sess = S.Session()
    x = S.Input(8, (32, 32, 3), slots=6)
    gan_model.gan(x, _args("iwgan", Lz=200))
    prm = sess.store.params
    want = {"discriminator/vars/c1/weights": ((5, 5, 3, 200), (5, 5, 3, 208)),
            "discriminator/vars/c1/bias": ((200,), (208,)),
            "discriminator/vars/c2/weights": ((5, 5, 200, 400), (5, 5, 208, 400)),
            "generator/vars/dc2/weights": ((5, 5, 200, 400), (5, 5, 208, 400)),
            "generator/BatchNorm_2/beta": ((200,), (208,)),
            "generator/vars/dc3/weights": ((5, 5, 3, 200), (5, 5, 3, 208)),
            "discriminator/vars/c3/weights": ((5, 5, 400, 800), (5, 5, 400, 800))}
    for name, (logical, physical) in want.items():
        assert prm[name].logical_shape == logical and prm[name].shape == physical, name
    # pad / slice round trip on host tensors
    t = torch.arange(2 * 3 * 5, dtype=torch.float32).reshape(2, 3, 5)
    padded = V._pad_to(t, (2, 4, 8))
    assert padded.shape == (2, 4, 8) and padded.sum() == t.sum() and torch.equal(padded[:, :3, :5], t)
    p = E.Param("x", (2, 4, 8), logical_shape=(2, 3, 5))
    assert torch.equal(p.logical(padded.reshape(-1)), t)
    # switch: no padding, physical == logical
    L.CHANNEL_PAD = False
    try:
        sess2 = S.Session()
        gan_model.gan(S.Input(8, (32, 32, 3), slots=6), _args("iwgan", Lz=200))
        assert all(q.shape == q.logical_shape for q in sess2.store.params.values())
    finally:
        L.CHANNEL_PAD = True


def test_deferred_update_runs_inline_without_a_device():
    """Session.defer_update is the multi-GPU overlap hook (gradient exchange on a side stream); during the
    graph-construction pass / without CUDA it must simply run the callable so that the schedule is unchanged."""
    sess = S.Session()
    calls = []
    prev = E.S.dry
    E.S.dry = True
    try:
        sess.defer_update(lambda: calls.append("exchange"))
        sess.join_updates()                       # nothing pending: no-op
    finally:
        E.S.dry = prev
    assert calls == ["exchange"] and sess._join_event is None
    sess.overlap_updates = False
    sess.defer_update(lambda: calls.append("inline"))
    assert calls == ["exchange", "inline"]


def test_latent_size_whose_channel_padding_meets_a_reshape():
    """latent_size=50: the critic's c3 output has 200 channels stored as 208; the reshape to rows of 4*4*4L
    (models/gan.py:284) must strip the padding, not mix zero channels into the rows (B=100 used to yield 104
    rows silently).  The reference accepts any latent_size."""
    sess = S.Session()
    x = S.Input(100, (32, 32, 3), slots=6)
    train = gan_model.gan(x, _args(B=100, Lz=50))
    E.S.dry = True
    try:
        sess.store.begin_pass()
        names = []
        orig = E.launch

        def rec(name, *a, **k):
            names.append(name)
            return 0
        E.launch = rec
        try:
            gl, dl = train.tower(x.next(), "d")
        finally:
            E.launch = orig
    finally:
        E.S.dry = False
    assert "b200_slice_cols" in names                      # the un-padding copy ran before each fc2
    t = E.Tensor(torch.empty((100, 4, 4, 208), dtype=torch.bfloat16, device="meta")); t.logical_c = 200
    E.S.dry = True
    try:
        assert E.reshape(t, (-1, 4 * 4 * 200)).shape == (100, 3200)
        with pytest.raises(b200gan._capi.B200Error):
            E.reshape(E.Tensor(torch.empty((100, 3200), device="meta")), (-1, 3000))
    finally:
        E.S.dry = False


def test_checkpoint_round_trip_by_tf_names(tmp_path):
    """3dgan_b200/checkpoint.py on host buffers: variables, TF-named optimizer slots (logical shapes, channel padding
    stripped), step counters and batch-norm moving averages survive save -> fresh session -> restore."""
    from b200gan import checkpoint as CK
    from b200gan.variables import optimizer_cfg

    def build():
        sess = S.Session()
        x = S.Input(4, (32, 32, 3), slots=6)
        gan_model.gan(x, _args("wgan", B=4, Lz=50))              # latent 50: 200-channel layers stored as 208
        st = sess.store
        st.finalize([("generator", st.collection("generator"), optimizer_cfg(_args())),
                     ("discriminator", st.collection("discriminator"), optimizer_cfg(_args()))], torch.device("cpu"))
        sess.counter = torch.zeros(1, dtype=torch.int64)
        return sess

    a = build()
    g = torch.Generator().manual_seed(3)
    for grp in a.store.groups:
        for p in grp.params:                                     # write only the logical block: padding stays zero
            p.logical(p.p32).copy_(torch.randn(p.logical_shape, generator=g))
            p.logical(grp.m[p.offset:p.offset + p.numel]).copy_(torch.randn(p.logical_shape, generator=g))
            p.logical(grp.v[p.offset:p.offset + p.numel]).copy_(torch.rand(p.logical_shape, generator=g))
        grp.step.fill_(7)
    for sv in a.store.state.values():
        sv.logical(sv.buf).copy_(torch.randn(sv.logical_shape, generator=g))
    a.counter.fill_(42)
    path = CK.save(a, str(tmp_path), 3, global_epoch=3, extra={"note": "x"})
    ck = torch.load(path, weights_only=False)
    assert ck["variables"]["discriminator/vars/c3/weights"].shape == (5, 5, 100, 200)
    assert ck["slots"]["discriminator/vars/c3/weights/Adam_1"].shape == (5, 5, 100, 200)
    assert ck["state"]["discriminator/BatchNorm_1/moving_variance"].shape == (200,)
    assert ck["global_step"] == 14 and ck["global_epoch"] == 3
    assert abs(ck["optimizers"]["generator"]["optimizers/beta1_power"] - 0.5 ** 8) < 1e-12
    b = build()
    assert CK.latest(str(tmp_path)) == path
    got = CK.restore_latest(b, str(tmp_path))
    assert got[1:3] == (14, 3) and got[3]["note"] == "x" and int(b.counter.item()) == 42
    for ga, gb in zip(a.store.groups, b.store.groups):
        assert torch.equal(ga.p32, gb.p32) and torch.equal(ga.m, gb.m) and torch.equal(ga.v, gb.v)
        assert int(gb.step.item()) == 7 and torch.equal(gb.p16, gb.p32.to(torch.bfloat16))
    assert torch.equal(a.store.state_buf, b.store.state_buf)
