"""Two ranks on two GPUs (torchrun, NCCL): the mu-block sharded solve of one large grid (BASELINE configs[3]) against the
unsharded solve on one GPU -- both exchange schemes: the NCCL all-gather of the I_n blocks (MuShardedSolver.solve) and the
fused one where the contraction kernel reads the peers' blocks by TMA over NVLink (solve_p2p, sos_source_peers, CUDA-IPC
ping-pong buffers).  A reader that ran ahead of a writer on the other GPU would show up here as a deviation: the sharded
result must equal the unsharded solve with the same (general) contraction kernel bit for bit, and the default (folded)
unsharded solve to rounding.  Skipped on boxes with fewer than two GPUs (run with `gpurun --gpus 2`).
"""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _two_gpus():
    import torch
    return torch.cuda.is_available() and torch.cuda.device_count() >= 2


@pytest.mark.parametrize("ref_general", [True, False])
@pytest.mark.parametrize("p2p", [False, True])
def test_mu_sharded_solve_on_two_gpus_equals_unsharded(p2p, ref_general):
    if not _two_gpus():
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29611 + int(p2p) + 2 * int(ref_general)), os.path.join(ROOT, "tools", "mu_shard_check.py"),
           "--layers", "1500", "--angles", "256", "--tau", "6.0", "--orders", "12", "--phase", "fwc", "--check"]
    if p2p:
        cmd.append("--p2p")
    if ref_general:
        cmd.append("--ref-general")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    rec = json.loads(line)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "mu_shard_2gpu_%s_%s.json" % ("p2p" if p2p else "nccl", "general" if ref_general else "folded")), "w") as f:
        f.write(line + "\n")
    assert rec["world"] == 2 and rec["p2p"] == p2p and rec["status"] == 0
    assert rec["orders"] == rec["orders_unsharded"] == 12
    if ref_general:
        assert rec["reference_contraction"] == "general" and rec["max_rel_dev_vs_unsharded"] == 0.0, rec
    else:
        assert rec["max_rel_dev_vs_unsharded"] < 1e-12, rec
