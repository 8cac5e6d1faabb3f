"""Multi-GPU check of the column-parallel BFP linear (SURVEY.md section 8 e): the fused path (all-gather in the GEMM epilogue
over peer memory) and the NCCL path must both equal the single-GPU BFPLinear bit for bit.  Needs >= 2 GPUs; skipped otherwise
(the host-side sharding logic is covered on CPU by tests/test_dist_cpu.py)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_column_parallel_fused_and_nccl_equal_single_gpu(tmp_path):
    out = tmp_path / "cp.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "tools", "check_column_parallel.py"), str(out), "--small"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rows = json.load(open(out))
    assert len(rows) == 2
    for row in rows:
        assert row["bit_equal_to_single_gpu"] and row["nccl_path_bit_equal"], row
    assert rows[0]["fused_path"] or rows[0]["fused_failed"], rows[0]      # fused unless symmetric memory is unavailable (then recorded)
