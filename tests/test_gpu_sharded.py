"""GPU parity of the row-sharded multi-GPU path (needs >= 2 GPUs: `gpurun --gpus 2`): W ranks with
row-sharded tables and NCCL all-to-alls must reproduce the single-GPU trainer on the global batch."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("mode", ["eager", "graph"])
def test_sharded_matches_single_gpu(mode):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29731", os.path.join(HERE, "sharded_worker.py"), mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    print(r.stdout[-4000:])
    print(r.stderr[-4000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "FAIL" not in r.stdout
