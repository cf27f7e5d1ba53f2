"""GPU, >= 2 devices: the data-parallel path on real NCCL — exchanged gradients and post-update parameters vs the
oracle's multi-tower `average_gradients` (util.py:118-147), with the overlapped and the in-line exchange."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("overlap", ["1", "0"])
def test_two_rank_nccl_step_matches_two_tower_oracle(tmp_path, overlap):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    out = str(tmp_path / "report.json")
    env = dict(os.environ, B200GAN_OVERLAP_UPDATE=overlap, B200GAN_BUCKET_MB="0")    # tiny model: one bucket per variable
    port = 29600 + (os.getpid() % 1000) + int(overlap)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py"), "--out", out]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    rep = json.load(open(out))
    print(rep)
    assert rep["world"] == 2 and rep["nccl_abi"] and rep["overlap"] == (overlap == "1")
    assert rep["buckets_d"] >= 4                      # the exchange really went out in several buckets
    assert rep["ok"], rep
