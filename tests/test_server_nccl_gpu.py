"""C5 on real ranks: torchrun, one process per GPU, NCCL over NVLink; both exchanges against the oracle ("both must give
identical top-2").  Needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_server_nccl_gpu.py -m gpu` (log in profiles/)."""
import os
import subprocess
import sys

import pytest

from multi_orbslam3_b200 import orbx

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c5_exchanges_on_nccl_ranks():
    n = orbx.lib().orbx_device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = 2 if n < 4 else 4
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", str(29600 + os.getpid() % 300), os.path.join(ROOT, "tests", "nccl_c5_worker.py")],
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, (p.stdout[-1500:], p.stderr[-3000:])
    assert "nccl c5 ok" in p.stdout
