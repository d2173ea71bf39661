"""Multi-GPU parity (needs >= 2 GPUs: `gpurun --gpus N -- python -m pytest tests/test_gpu_multi.py -m gpu`, N = 2, 4 or 8): row
sharding with the native ncclAllReduce exchange (north_star config c5's mode) and chain sharding (c1-c4's mode), against a
single-GPU run of the same chains.  Every world size the box offers is run; the others are skipped."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_row_and_chain_sharding(world):
    import torch
    n = torch.cuda.device_count()
    if n < world:
        pytest.skip(f"needs {world} GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", str(29531 + world), os.path.join(HERE, "multi_gpu_worker.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and f"MULTI_GPU_OK world={world}" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
