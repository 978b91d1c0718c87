"""N-GPU parity as a pytest case: spawns tests/multigpu_check.py under torch.distributed.run when the box has >= 2 GPUs
(every owner compares its tables with the oracle on the union of all ranks' reads, peer-memory and NCCL exchange)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("transport", ["peer", "nccl"])
def test_sharded_build_matches_oracle(transport):
    import torch

    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run by `gpurun --gpus N`; bench.py --gpus N carries the same check in its `parity` object)")
    world = 2 if n < 4 else (4 if n < 8 else 8)
    env = dict(os.environ, GGCAT_B200_EXCHANGE=transport)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", "29517", str(ROOT / "tests" / "multigpu_check.py")],
                         capture_output=True, text=True, timeout=600, cwd=str(ROOT), env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert "multigpu_check ok" in out.stdout
