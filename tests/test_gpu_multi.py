"""Multi-GPU (one process per GPU, NCCL) check — runs only where >= 2 GPUs are visible."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_tensor_parallel_and_data_parallel_on_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (covered on CPU by tests/test_parallel_cpu.py)")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = os.path.join(os.path.dirname(__file__), "mgpu_tp_check.py")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), script],
                         capture_output=True, text=True, timeout=600)
    assert "TP_CHECK_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
