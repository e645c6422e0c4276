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


@pytest.mark.parametrize("case", [
    # (layers, angles, tau*, order cap, chunk rows, folded contraction): thick (the widest extrapolation class, no surviving
    # windowed column), thin (windowed |mu| < 0.01 columns survive: tau-window halos of hundreds of rows across the block
    # boundary), and the general (unfolded) contraction kernel on its row-restricted path
    (1500, 256, 6.0, 300, 64, True),
    (1200, 512, 0.05, 300, 48, True),
    (1100, 128, 2.0, 300, 0, False),
])
def test_layer_sharded_solve_is_bit_identical_to_unsharded(case):
    """Layer-block sharding (csrc/layer_shard.cuh): two ranks exchange chunk aggregates, halo rows and ratios by stores into
    each other's memory from inside the CUDA-graphed order loop.  With fewer than two GPUs the two ranks share GPU 0
    (time-sliced: slow, but the same code path), so the driver's one-GPU run covers it too."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    L, M, tau, cap, chunk, fold = case
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29631 + (L % 7)), os.path.join(ROOT, "tools", "layer_shard_check.py"),
           "--layers", str(L), "--angles", str(M), "--tau", str(tau), "--orders", str(cap), "--chunk-rows", str(chunk), "--repeat", "1", "--check"]
    env = dict(os.environ, SOS_B200_FOLD="1" if fold else "0")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    rec = json.loads(line)
    assert rec["folded"] == fold, rec
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "layer_shard_2rank_L%d_M%d.json" % (L, M)), "w") as f:
        f.write(line + "\n")
    assert rec["world"] == 2 and rec["status"] == 0, rec
    assert rec["orders"] == rec["orders_unsharded"], rec
    assert rec["bit_identical"] and rec["max_rel_dev_vs_unsharded"] == 0.0, rec
