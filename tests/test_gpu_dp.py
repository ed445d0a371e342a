"""Two-GPU data-parallel checks (skipped on a single-GPU box): gradient buckets overlapped with the backward walk are
bit-identical to one all-reduce per net, graph replay is identical, and sync_bn over 2 shards equals the single-GPU step on
the concatenated batch (scripts/dp_check.py)."""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_bucket_overlap_and_sync_bn():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "dp_check.py")]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600).stdout
    lines = [ln for ln in out.splitlines() if ln.startswith("[")]
    assert len(lines) == 3, out[-2000:]
    for ln in lines[:2]:
        assert "G identical True, D identical True" in ln and "graph replay identical True" in ln, ln
    errs = [float(x) for x in re.findall(r"rel err G ([0-9.e+-]+) D ([0-9.e+-]+)", lines[2])[0]]
    assert max(errs) <= 1e-5, lines[2]


def test_two_gpu_peer_memory_allreduce_of_sync_bn_statistics():
    """The one-shot NVLink peer-memory all-reduce fused into the BatchNorm statistics tail (kernels_peer.cu) against
    ncclAllReduce of the same sums: two terms have one summation order, so the parameters after three steps are bit-identical,
    eagerly and under graph replay (the call counter lives in device memory)."""
    import json

    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29534", os.path.join(ROOT, "scripts", "peer_ar_check.py")]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=900).stdout
    res = json.loads([ln for ln in out.splitlines() if ln.startswith("{")][-1])
    assert res["sync_bn_peer_allreduce"] is True, res
    assert res["sync_bn_peer_vs_nccl_max_rel_diff"] == 0.0, res
    assert res["sync_bn_peer_graph_replay_identical"] is True, res
    assert res["sync_bn_rel_err"] <= 1e-5, res
