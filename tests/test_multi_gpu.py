"""The path's single collective on real GPUs: the library-owned NCCL communicator (pimdk_comm_init / pimdk_ti_allreduce /
pimdk_ti_reduce_dev) across two ranks of one box.  Needs >= 2 GPUs; skipped otherwise (the world-size-2 gloo rig of
tests/test_host.py covers the host logic on CPU)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
def test_library_owned_nccl_allreduce_two_ranks():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(HERE, "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "mgpu ok: 2 ranks" in r.stdout
