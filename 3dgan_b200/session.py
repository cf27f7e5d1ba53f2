"""Session: what `tf.Session` + `tf.train.Supervisor` meant to the hot path (train.py:254-308) —
device binding, the variable store, noise streams, the input feed, CUDA-graph capture of a step and
the data-parallel exchange (one process per GPU; NCCL all-reduce of the flat gradient bucket in
place of util.py:118-147's CPU averaging).
"""
import contextlib
import gc
import os

import torch

from . import engine as E
from .ops import layers as L
from .variables import VariableStore

_current = None


def current():
    if _current is None:
        raise RuntimeError("no b200gan Session is active")
    return _current


class Input:
    """The `x` tensor the reference's input pipeline hands to a model (data.py:34-60): yields one
    [B,H,W,C] float32 batch in [0,1] per `next()`.  Batches are fed by the caller (`feed`) into a
    static device ring so that a captured CUDA graph always reads the same addresses."""

    def __init__(self, batch_size, shape, slots=1, device=None, dtype=torch.float32, layout="NHWC"):
        """dtype float32: batches already normalised to [0,1] (what the reference's `x` tensor holds);
        dtype uint8: image bytes as decoded — the /255 of data.py:21-22,29 then runs on the device, fused into the
        model's first op (engine.affine), and a batch costs a quarter of the host->device traffic.
        layout "NCHW": `shape` is (C, H, W), batches arrive as the reference's Gen-2 pipeline lays them out
        (hem/ops/layers.py:117-119); the model's first op (hem.rescale) transposes them on the device."""
        self.batch_size, self.shape, self.slots = batch_size, tuple(shape), slots
        self.device = device
        self.dtype = dtype
        self.layout = layout
        self.ring = None
        self.cursor = 0

    def materialize(self, device):
        self.device = device
        self.ring = torch.zeros((self.slots, self.batch_size) + self.shape, dtype=self.dtype, device=device)

    def feed(self, slot, host_or_device_batch, non_blocking=True):
        self.ring[slot].copy_(host_or_device_batch, non_blocking=non_blocking)

    # ---- prefetching feed (the role of the reference's queue runners, data.py:34-60): the next iteration's
    # batches travel host -> device on a copy stream while the current iteration computes
    def prefetch(self, pinned_host_batches):
        """Start the asynchronous copy of one iteration's batches ([slots, B, H, W, C] pinned host tensor)
        into the staging buffer.  Call `commit()` before the iteration that consumes them."""
        if getattr(self, "_staging", None) is None:
            self._staging = torch.empty_like(self.ring)
            self._copy_stream = torch.cuda.Stream(device=self.ring.device)
            self._staged = torch.cuda.Event()
            self._consumed = torch.cuda.Event()
            self._consumed.record()
        self._copy_stream.wait_event(self._consumed)           # the previous commit has read the staging buffer
        with torch.cuda.stream(self._copy_stream):
            self._staging.copy_(pinned_host_batches, non_blocking=True)
            self._staged.record()

    def commit(self):
        """Make the prefetched batches the ring's contents (device-to-device, on the compute stream)."""
        cur = torch.cuda.current_stream()
        cur.wait_event(self._staged)
        self.ring.copy_(self._staging, non_blocking=True)
        self._consumed.record(cur)

    def reset(self):
        self.cursor = 0

    def next(self):
        if E.S.dry:
            t = E.Tensor(torch.empty((self.batch_size,) + self.shape, dtype=self.dtype, device="meta"))
        else:
            t = E.Tensor(self.ring[self.cursor % self.slots])
            self.cursor += 1
        t.layout = self.layout
        return t


class Session:
    def __init__(self, seed=0, noise_seed=1234, device=None):
        global _current
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.cuda = torch.cuda.is_available()
        if device is None:
            device = torch.device("cuda", self.local_rank) if self.cuda else torch.device("meta")
        self.device = device
        if self.cuda:
            torch.cuda.set_device(device)
        self.store = VariableStore(seed)
        L.set_store(self.store)
        self.noise_seed = noise_seed + self.rank        # per-tower noise
        self.counter = None
        self.noise_queue = []                           # parity tests inject noise here
        self.graphs = {}
        self.use_graphs = os.environ.get("B200GAN_CUDA_GRAPHS", "1") != "0"
        self.dist = None
        self.comm = None                                # b200_nccl_* communicator (ctypes void*), data plane
        self.muted = False                              # True: no collectives (rank-local profiling passes)
        self._exchanges = {}
        # gradient exchange + optimizer update of one run overlap the start of the next run (side stream)
        self.overlap_updates = os.environ.get("B200GAN_OVERLAP_UPDATE", "1") != "0"
        self._side = None
        self._join_event = None
        _current = self

    # ---------------------------------------------------------------- build / run modes
    @contextlib.contextmanager
    def building(self):
        """Graph-construction pass: shapes + variable creation only (works without a GPU)."""
        prev = E.S.dry
        E.S.dry = True
        self.store.begin_pass()
        try:
            yield
        finally:
            E.S.dry = prev

    def begin_step(self):
        if not self.cuda:
            raise RuntimeError("b200gan needs a CUDA device: there is no CPU fallback for the training step")
        E.begin(self.device)
        self.store.begin_pass()
        if self.counter is None:
            self.counter = torch.zeros(1, dtype=torch.int64, device=self.device)

    # ---------------------------------------------------------------- noise
    def _noise(self, shape, normal, stream_id, f32):
        if self.noise_queue and not E.S.dry:
            v = self.noise_queue.pop(0)
            assert tuple(v.shape) == tuple(shape), (tuple(v.shape), tuple(shape))
            return E.Tensor(v.to(self.device, torch.float32 if f32 else torch.bfloat16).contiguous())
        return E.random_fill(shape, normal, self.noise_seed, self.counter, stream_id, f32)

    def random_normal(self, shape, stream_id=0, f32=False):
        return self._noise(shape, True, stream_id, f32)

    def random_uniform(self, shape, stream_id=1, f32=True):
        return self._noise(shape, False, stream_id, f32)

    # ---------------------------------------------------------------- data parallel
    def init_distributed(self, backend="nccl"):
        """One process per GPU = one tower (util.py:54-77).  torch.distributed is the control plane (rendezvous,
        barriers, the bench's max-over-ranks); the gradient exchange itself goes through the C ABI's b200_nccl_*
        entry points on a communicator of our own (CPU-only runs / the gloo tests fall back to torch's all_reduce)."""
        import torch.distributed as dist
        if self.world > 1 and not dist.is_initialized():
            dist.init_process_group(backend=backend, rank=self.rank, world_size=self.world,
                                    device_id=self.device if backend == "nccl" else None)
        self.dist = dist if self.world > 1 else None
        if self.dist is not None and self.cuda and os.environ.get("B200GAN_NCCL_ABI", "1") != "0":
            import ctypes
            from . import _capi as K
            path = K.nccl_library_path()
            K.call("b200_nccl_load", path.encode() if path else None)
            uid = ctypes.create_string_buffer(128)
            if self.rank == 0:
                K.call("b200_nccl_unique_id", uid)
            t = torch.tensor(list(uid.raw), dtype=torch.uint8, device=self.device if backend == "nccl" else "cpu")
            dist.broadcast(t, 0)
            raw = bytes(t.cpu().tolist())
            comm = ctypes.c_void_p()
            K.call("b200_nccl_init", raw, self.rank, self.world, ctypes.byref(comm))
            self.comm = comm

    def all_reduce(self, flat):
        """In-place sum over ranks of a contiguous fp32 slice, asynchronous on the engine's current stream."""
        if self.muted:
            return
        if self.comm is not None:
            E.launch("b200_nccl_allreduce_f32", self.comm, E._p(flat), flat.numel())
        elif self.dist is not None:
            self.dist.all_reduce(flat)

    def broadcast(self, flat, src):
        if self.comm is not None:
            E.launch("b200_nccl_broadcast_f32", self.comm, E._p(flat), flat.numel(), src)
        elif self.dist is not None:
            self.dist.broadcast(flat, src)

    def all_reduce_grads(self, group):
        """average_gradients (util.py:118-147): sum over ranks here, the 1/n is folded into the
        optimizer kernel's grad_scale."""
        self.all_reduce(group.g32)
        return 1.0 / self.world

    def exchange(self, group):
        """The bucketed, overlapped form of all_reduce_grads for one optimizer group (one Exchange per group)."""
        ex = self._exchanges.get(id(group))
        if ex is None:
            ex = self._exchanges[id(group)] = Exchange(self, group)
        return ex

    # ---------------------------------------------------------------- overlapped parameter updates
    def side_stream(self):
        if self._side is None:
            # B200GAN_SIDE_PRIORITY=1: a high-priority stream, so that the exchange's CTAs are placed as soon as SMs
            # free up instead of queueing behind the next GEMM's grid
            prio = -1 if os.environ.get("B200GAN_SIDE_PRIORITY", "0") != "0" else 0
            self._side = torch.cuda.Stream(device=self.device, priority=prio)
        return self._side

    def defer_update(self, fn):
        """Run `fn` (the gradient all-reduce of one optimizer group: it touches only that group's gradient
        bucket and allocates nothing) on a side stream, ordered after everything issued so far.  The caller
        goes on with work that does not touch that bucket and calls `join_updates()` before the update.
        Inside a CUDA-graph capture the fork / join become graph edges, so the exchange runs concurrently
        with the next run's generator forward.  (Also running the optimizer kernel there was measured
        slower on one GPU: its blocks crowd the GEMM CTAs out of the SMs.)"""
        if not self.overlap_updates or E.S.dry:
            fn()
            return
        import ctypes
        side = self.side_stream()
        main = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(main)
        side.wait_event(fork)
        prev = E.S.stream
        E.S.stream = ctypes.c_void_p(side.cuda_stream)
        try:
            with torch.cuda.stream(side):
                fn()
                self._join_event = torch.cuda.Event()
                self._join_event.record(side)
        finally:
            E.S.stream = prev

    def join_updates(self):
        if self._join_event is not None:
            torch.cuda.current_stream().wait_event(self._join_event)
            self._join_event = None

    # ---------------------------------------------------------------- CUDA graphs
    def run(self, key, fn):
        """Run `fn()` (a whole step: forward, backward, exchange, update).  First call: eager
        (warm-up, lazy allocations).  Second call: captured into a CUDA graph.  Later: replay.
        fn returns a dict of fp32 [1] device tensors (losses)."""
        if not self.use_graphs or self.noise_queue:
            self.begin_step()
            return fn()
        ent = self.graphs.get(key)
        if ent is None:
            self.begin_step()
            out = fn()
            self.graphs[key] = {"state": "warm"}
            return out
        if ent["state"] == "warm":
            # Python's cyclic collector must not run inside the capture: if it finalises a CUDA graph or tensors of
            # an earlier Session there, the allocator's cudaFree invalidates the capture ("operation failed due to
            # a previous error during capture", seen intermittently when several Sessions live in one process)
            gc.collect()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            launches0 = E.S.launches
            gc_on = gc.isenabled()
            gc.disable()
            try:
                with torch.cuda.graph(g):
                    self.begin_step()
                    out = fn()
            finally:
                if gc_on:
                    gc.enable()
            ent.update(state="graph", graph=g, out=out, launches=E.S.launches - launches0)
            # the engine was bound to the capture stream: rebind it to the current stream so that eager launches
            # after this point (store.load -> refresh_transposed, tests) stay ordered with torch ops
            E.begin(self.device)
            # capture does not execute: replay once so this call has the step's effect
            g.replay()
            return out
        ent["graph"].replay()
        E.S.launches += ent["launches"]
        return ent["out"]


class Exchange:
    """average_gradients (util.py:118-147) for one optimizer group, overlapped with the backward pass.

    The group's flat fp32 gradient bucket is cut into buckets of >= `bucket_bytes` walking the variables in REVERSE
    creation order — the order in which a reverse sweep finishes them (the last layers first).  `on_ready(param)`
    (engine.backward calls it the moment a variable's gradient is final) starts a bucket's all-reduce on a side
    stream as soon as all of its variables are final, so the exchange of the deep layers (IWGAN critic: c3 + fc2 =
    33 MB of the 40 MB) runs under the remaining backward of the shallow ones; `finish()` sends what is left and
    `join()` makes the main stream wait for the side stream before the optimizer reads the bucket.  With one rank
    everything is a no-op.  Inside a CUDA-graph capture the fork / join events become graph edges."""

    def __init__(self, sess, group, bucket_bytes=None):
        self.sess, self.group = sess, group
        if bucket_bytes is None:
            bucket_bytes = int(os.environ.get("B200GAN_BUCKET_MB", "8")) << 20
        self.buckets = []                        # [lo, hi, set(param ids)] element ranges of the flat bucket
        hi, ids, lo = group.size, set(), group.size
        for p in reversed(group.params):
            lo = p.offset
            ids.add(id(p))
            if (hi - lo) * 4 >= bucket_bytes:
                self.buckets.append((lo, hi, ids))
                hi, ids = lo, set()
        if ids:
            self.buckets.append((0, hi, ids))
        self.begin()

    @property
    def active(self):
        s = self.sess
        return (s.dist is not None or s.comm is not None) and not s.muted

    def begin(self):
        self.ready, self.next = set(), 0
        self._event = None

    def on_ready(self, param):
        if not self.active or E.S.dry:
            return
        self.ready.add(id(param))
        while self.next < len(self.buckets) and self.buckets[self.next][2] <= self.ready:
            self._fire(self.buckets[self.next])
            self.next += 1

    def finish(self):
        """Send every bucket that has not gone yet (variables the sweep never reported included)."""
        if not self.active or E.S.dry:
            return
        while self.next < len(self.buckets):
            self._fire(self.buckets[self.next])
            self.next += 1

    def _fire(self, bucket):
        import ctypes
        lo, hi, _ = bucket
        sess = self.sess
        if not sess.cuda or not sess.overlap_updates:
            sess.all_reduce(self.group.g32[lo:hi])
            return
        side, main = sess.side_stream(), torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(main)
        side.wait_event(fork)
        prev = E.S.stream
        E.S.stream = ctypes.c_void_p(side.cuda_stream)
        try:
            with torch.cuda.stream(side):
                sess.all_reduce(self.group.g32[lo:hi])
                self._event = torch.cuda.Event()
                self._event.record(side)
        finally:
            E.S.stream = prev

    def join(self):
        """Main stream waits for the exchange; returns the optimizer's grad_scale (the 1/n of the average)."""
        if self._event is not None:
            torch.cuda.current_stream().wait_event(self._event)
            self._event = None
        return 1.0 / self.sess.world
