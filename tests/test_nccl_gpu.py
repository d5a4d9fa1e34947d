"""GPU, >= 2 devices: results produced by SEVERAL physical GPUs over real NCCL all-gathers (one process per
GPU, launched like the bench) against the CPU oracle and against the unsharded single-GPU run.  Skipped
on a one-GPU box (the one-process LocalGroup tests cover the protocol there); `gpurun --gpus 2` runs it."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world, kind, nv, nseg):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "nccl_worker.py"), kind, str(nv), str(nseg)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    outs = [json.loads(ln.split("NCCLWORKER ", 1)[1]) for ln in r.stdout.splitlines() if "NCCLWORKER " in ln]
    assert len(outs) == world
    return sorted(outs, key=lambda o: o["rank"])


@pytest.mark.parametrize("kind,nv,nseg", [("c2", 24, 600), ("c4", 12, 3000)])
def test_real_nccl_ranks_match_the_oracle_and_the_single_gpu_run(kind, nv, nseg):
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    outs = _run(world, kind, nv, nseg)
    ref = outs[0]["digest_unsharded"]
    for o in outs:
        assert o["digest_step0"] == o["digest_step1"] == o["digest_step2"] == ref, o
        assert o["sizes"]["entries"] > 0
