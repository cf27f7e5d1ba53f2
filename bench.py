#!/usr/bin/env python
"""bench.py — IWGAN 32x32x3 training throughput (BASELINE.json metric) on N B200s of one node.

A "step" is one `train_func` call of the reference (train.py:307): n_disc_train(5) critic updates + 1
generator update (+ the d_loss report), each on a fresh batch of 512 images per GPU
(BASELINE configs[1]; weak scaling = the reference's per-tower batch semantics).

  python bench.py [--gpus N --steps K --warmup W]            this repo's CUDA path
  python bench.py --impl reference [...]                     CPU restatement of the reference (oracle)

Prints ONE JSON line (see the task contract): value = images/s with inputs resident in HBM, device
timed, max over ranks; e2e = same through the public API with pinned-host batches copied in (Input.prefetch /
commit: the next step's H2D copy overlaps this step's compute) and losses read back every step; roofline = the dominant kernel family (tcgen05 implicit GEMM) timed live with
CUDA events per launch; cpu_baseline = the oracle on this box's host cores (bounded sample).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "IWGAN 32x32x3 train images/sec"
UNIT = "images/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--latent", type=int, default=200)
    ap.add_argument("--size", type=int, default=32)
    ap.add_argument("--n_disc_train", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--profile-out", default=None, help="write the per-launch GEMM table (json) here")
    return ap.parse_args()


def workload(a):
    return {"workload": "iwgan_%dx%dx3_b%d_L%d (BASELINE configs[1]: 5 critic + 1 generator update per step)"
                        % (a.size, a.size, a.batch, a.latent),
            "batch_per_gpu": a.batch, "latent_size": a.latent, "n_disc_train": a.n_disc_train,
            "optimizer": "adam lr=1e-4 beta=(0.5,0.9)"}


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], False
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for nme, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- reference arm
def cpu_reference(a, steps, warmup, sample_batch):
    """The reference's CPU path restated (oracle.models.GanTrainer): same graph, schedule, optimizer and
    synthetic inputs; one tower; all host threads.  Each step is a bounded sample (a smaller batch)."""
    import torch
    from oracle import models as OM
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tr = OM.GanTrainer("iwgan", a.size, 3, a.latent, sample_batch, lr=1e-4, beta1=0.5, beta2=0.9,
                       n_disc=a.n_disc_train, seed=0)
    gen = torch.Generator().manual_seed(1234)
    nb = lambda: torch.rand(sample_batch, a.size, a.size, 3, generator=gen)
    nn = lambda: (torch.randn(sample_batch, a.latent, generator=gen), torch.rand(sample_batch, 1, generator=gen))
    for _ in range(warmup):
        tr.iteration(nb, nn)
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.iteration(nb, nn)
    dt = time.perf_counter() - t0
    return sample_batch * steps / dt, dt / steps * 1e3, cores


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 32
    v, ms, cores = cpu_reference(a, a.steps, a.warmup, sample)
    cfg = workload(a)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "every step = one full iteration (5 critic + 1 generator) at batch %d "
                                       "instead of %d; torch-CPU restatement of the TF graph (TF not installable)"
                                       % (sample, a.batch)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------- our arm
def run_b200(a):
    import torch
    import b200gan  # noqa: F401
    from b200gan import engine as E
    from b200gan import session as S
    from b200gan.models import gan as gan_model

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the b200 arm has no CPU fallback")
    verbose = bool(os.environ.get("B200GAN_BENCH_VERBOSE"))

    def mark(msg):
        if verbose:
            print("[bench rank %s] %s" % (os.environ.get("RANK", "0"), msg), file=sys.stderr, flush=True)
    sess = S.Session(seed=0, noise_seed=1234)
    world, rank = sess.world, sess.rank
    if world > 1:
        sess.init_distributed("nccl")
    args = argparse.Namespace(model="iwgan", batch_size=a.batch, latent_size=a.latent, n_disc_train=a.n_disc_train,
                              optimizer="adam", lr=1e-4, beta1=0.5, beta2=0.9)
    runs = a.n_disc_train + 1
    x = S.Input(a.batch, (a.size, a.size, 3), slots=runs)
    train = gan_model.gan(x, args)
    mark("model built")

    # synthetic data: two sets of `runs` device-resident batches (fresh batch per sess.run-equivalent)
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    pool = torch.rand((2, runs, a.batch, a.size, a.size, 3), generator=gen, device="cuda")
    host = torch.rand((2, runs, a.batch, a.size, a.size, 3)).pin_memory()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def step_resident(i):
        x.ring.copy_(pool[i & 1])
        return sess.run("gan_iteration", train.iteration)

    def step_e2e(i):
        # the public feed API: this step's batches were prefetched (pinned host -> device on a copy stream)
        # during the previous step; the next step's copy starts before this step's losses are read back
        x.commit()
        out = sess.run("gan_iteration", train.iteration)
        if i + 1 < a.steps:
            x.prefetch(host[(i + 1) & 1])
        return {k: float(v.item()) for k, v in out.items()}

    for i in range(max(a.warmup, 3)):
        out = step_resident(i)
        if verbose:
            mark("warmup step %d %s" % (i, {k: float(v.item()) for k, v in out.items()}))
    barrier()
    mark("warm-up done")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = E.S.launches
    clk = ClockSampler(sess.local_rank)
    clk.__enter__()
    barrier()
    ev0.record()
    for i in range(a.steps):
        step_resident(i)
    ev1.record()
    barrier()
    launches = E.S.launches - launches0
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t.item())
    value = a.batch * world * a.steps / (ms * 1e-3)

    # ---- end to end: pinned host batches in, losses out, every step
    barrier()
    t0 = time.perf_counter()
    last = None
    x.prefetch(host[0])                 # inside the timed region: every step's H2D copy is timed
    for i in range(a.steps):
        last = step_e2e(i)
        if os.environ.get("B200GAN_BENCH_VERBOSE"):
            print("e2e step", i, last, file=sys.stderr, flush=True)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    e2e_value = a.batch * world * a.steps / float(t.item())
    clk.__exit__(None, None, None)
    h2d = runs * a.batch * a.size * a.size * 3 * 4
    d2h = 8

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": dict(workload(a), l2="working set per step (activations+weights, >1 GB) exceeds the 126 MB L2",
                           parallelism="dp%d" % world, cuda_graph=bool(sess.use_graphs)),
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "losses": last}

    if rank == 0 and not a.no_roofline:
        line["roofline"] = roofline(sess, train, x, pool, a, line["ms_per_step"])
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        v, _, cores = cpu_reference(a, 2, 1, 32)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "2 timed iterations (after 1 warm-up) at batch 32 instead of %d, "
                                          "torch-CPU restatement of the reference graph" % a.batch}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # captured graphs hold NCCL work: drop them before tearing the communicator down, and do not
        # linger in interpreter shutdown (destroy_process_group can wait forever on captured collectives)
        torch.cuda.synchronize()
        torch.distributed.barrier()
        sess.graphs.clear()
        torch.cuda.synchronize()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


def roofline(sess, train, x, pool, a, step_ms):
    """Dominant kernel family = the tcgen05 implicit-GEMM launches (conv fprop/dgrad/wgrad).  One extra
    step is run eagerly with a CUDA-event pair around every such launch on the launching stream;
    achieved = sum(algorithmic FLOPs) / sum(durations) over those launches."""
    import torch
    from b200gan import engine as E
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, src = 1590.0, "fallback"
    if os.path.exists(peaks_path):
        pk = json.load(open(peaks_path))
        peak, src = float(pk.get("bf16_tflops_sustained", pk.get("bf16_tflops", 1590.0))), "measured (sustained: timed inside a long step)"
    prev, prev_dist = sess.use_graphs, sess.dist
    sess.use_graphs, sess.dist = False, None        # rank-0-only pass: no collective inside
    E.S.profile = []
    x.ring.copy_(pool[0])
    it0, it1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # eager launches are host-bound: park the GPU behind a ~0.2 s spin kernel so the whole iteration is
    # queued before it starts and the per-launch events see back-to-back execution (as in the graph replay)
    torch.cuda._sleep(int(4e8))
    it0.record()
    sess.run("profile", train.iteration)
    it1.record()
    torch.cuda.synchronize()
    iter_ms = it0.elapsed_time(it1)
    recs, E.S.profile = E.S.profile, None
    sess.use_graphs, sess.dist = prev, prev_dist
    rows = {}
    tot_f = tot_ms = 0.0
    for name, flops, e0, e1, tag in recs:
        ms = e0.elapsed_time(e1)
        if flops <= 0:
            continue
        r = rows.setdefault(tag, {"launches": 0, "flops": 0.0, "ms": 0.0})
        r["launches"] += 1; r["flops"] += flops; r["ms"] += ms
        if tag.startswith("tc:"):
            tot_f += flops; tot_ms += ms
    for r in rows.values():
        r["tflops"] = r["flops"] / (r["ms"] * 1e-3) / 1e12 if r["ms"] > 0 else None
    if a.profile_out:
        json.dump(rows, open(a.profile_out, "w"), indent=1)
    n_tc = sum(r["launches"] for t_, r in rows.items() if t_.startswith("tc:"))
    achieved = tot_f / (tot_ms * 1e-3) / 1e12 if tot_ms > 0 else 0.0
    return {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": ncu_traffic_bytes(), "peak_source": src,
            "kernel": "tapgemm2sm_kernel+wgrad2sm_kernel (tcgen05 cta_group::2 implicit GEMM: conv fprop/dgrad/wgrad, dense)",
            "launches": n_tc, "flops_per_launch_avg": tot_f / max(n_tc, 1), "ms_per_launch_avg": tot_ms / max(n_tc, 1),
            "step_share": tot_ms / step_ms if step_ms > 0 else None,
            "note": "achieved = sum(2*N*Ho*Wo*k*k*Cin*Cout over the launches) / sum(CUDA-event durations), one eager "
                    "iteration; step_share = those durations / the graph-replayed ms_per_step; traffic = mean dram read+write bytes per launch of the profiled c2/c3 launches "
                    "(profiles/r1_ncu_kernels.csv)"}


def ncu_traffic_bytes():
    """Mean DRAM bytes (read+write) per launch of the tcgen05 kernels in the committed ncu --set full capture."""
    import csv
    path = os.path.join(ROOT, "profiles", "r1_ncu_kernels.csv")
    if not os.path.exists(path):
        return None
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    try:
        ri, wi = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    except ValueError:
        return None
    vals = [(float(r[ri]) + float(r[wi])) * 1e6 for r in rows[3:] if r and ("tapgemm" in r[0] or "wgrad" in r[0])
            and float(r[hdr.index("gpu__time_duration.sum")]) > 100.0]
    return sum(vals) / len(vals) if vals else None


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
