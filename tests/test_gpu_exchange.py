"""Fused in-kernel gradient exchange (gaussianip_b200/exchange.py, csrc/preprocess_bwd.cu accumulate modes 2/3)
against the NCCL all-reduce of the bucket.  Needs two GPUs: runs scripts/exchange_check.py under torchrun."""
import ast
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("algo,extra", [("push_all", ["--sh", "1"]), ("owner_push", ["--sh", "1"]),
                                        ("owner_push", ["--precomp"]), ("push_all", ["--precomp", "--points", "50001"])])
def test_fused_exchange_equals_nccl_all_reduce(algo, extra):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "scripts", "exchange_check.py"), "--algo", algo,
           "--points", "60000", "--res", "256", "--views", "2", "--steps", "3"] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if not lines:
        assert "NVLS multicast is not available" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
        pytest.skip("fused exchange unavailable on this box: " + r.stdout[-300:])
    d = ast.literal_eval(lines[-1])
    assert d["radii_equal"] and d["replica_checksum_spread"] == 0.0
    assert d["rel"] <= 1e-5, d          # fp32 sums in a different order
    assert abs(d["loss_nccl"] - d["loss_fused"]) <= 1e-6 * abs(d["loss_nccl"])
