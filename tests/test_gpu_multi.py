"""Multi-GPU path on real devices (needs >= 2 GPUs; skipped otherwise): one process per GPU under torchrun,
protein-sharded index, ks_shard_search_batch (NCCL exchange + counting merge on rank 0) compared with the oracle for the
general, the dense and a scaled > 1 configuration.  Logs of runs on 2 and 8 B200s are committed under profiles/."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs at least 2 GPUs")
def test_sharded_search_matches_oracle():
    n = min(_n_gpus(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", "29613", os.path.join(ROOT, "tests", "_multi_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "all ok=True" in r.stdout, r.stdout[-2000:]
    print(r.stdout)
