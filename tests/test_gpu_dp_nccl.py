"""N-rank NCCL run of the REAL kernels == the single-GPU run of the concatenated batch (VERDICT r01:
"nothing checks that an NCCL 2-rank run equals the single-GPU big batch on hardware").  Spawns
``torch.distributed.run`` with one rank per visible GPU (2, or all of them up to 8); skipped on a single-GPU box.
Covers the resident fp32 / bf16 kernels and the streaming kernels, the Philox stream keyed by the GLOBAL chain index,
the single all-reduce that carries the weight gradients and the per-step scalars, and the global-batch normalisation."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs on the box")
def test_nccl_sharded_run_equals_single_gpu_big_batch():
    n = min(torch.cuda.device_count(), 8)
    n = 2 if n < 4 else (4 if n < 8 else 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(HERE, "dp_nccl_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("DP_NCCL_REPORT ")]
    assert line, (out.stdout[-2000:], out.stderr[-4000:])
    rep = json.loads(line[-1][len("DP_NCCL_REPORT "):])
    print(json.dumps(rep, indent=1)[:4000])
    assert rep["world"] == n
    assert rep["ok"], rep
    assert out.returncode == 0
