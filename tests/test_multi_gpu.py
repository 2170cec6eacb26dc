"""Two real devices (run with -m gpu on a box with >= 2 GPUs; skipped otherwise).

configs[4]: the 64x64-column terrain world is cut into two chunk-column ranges, one process per
GPU under torchrun (NCCL); every owned lightmap + mesh is hashed on its rank, the digests are
all-gathered and rank 0 compares them with a single-GPU build of the whole world made in the
same run (bench.py --scaling strong).  The single-device emulation of the same split is
tests/test_gpu_scale.py::test_config4_sharded_equals_single.
"""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_device_count() < 2, reason="needs two GPUs")
def test_two_ranks_equal_single_gpu_build():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.join(ROOT, "bench.py"),
           "--gpus", "2", "--scaling", "strong", "--steps", "2", "--warmup", "1", "--sustain", "0",
           "--copy-ceiling", "0"]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    res = json.loads(line)
    assert res["n_gpus"] == 2 and res["scaling"] == "strong"
    par = res["parity"]
    assert par["chunks"] == 64 * 64 and par["gathered"] == 64 * 64
    assert par["mismatches"] == 0, par
