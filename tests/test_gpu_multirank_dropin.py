"""GPU (needs >= 2 devices; skipped on a one-GPU box): the drop-in sweep on 2 GPUs - one process per GPU under torchrun, the operator
terms of multiplyH / diagonalH and the noise operators partitioned over the ranks (distribute.C's boost::mpi split -> b2d_plan(rank,
nranks)), partial sigma / diag(H) / noise density matrices all-reduced over NCCL inside the library; the eigen-decomposition of the density matrix
(sectors) and the operator rotation (operators) are divided over the ranks as well (option partition_renormalisation).  Every rank must print the same sweep
energies, and they must be the unmodified reference's (same bounds as tests/test_gpu_dropin.py)."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c2_d2h_M50_noise", "hubbard_L16_M80", "hubbard_L16_M1000"])   # the last one has sectors for the (partitioned) block-Jacobi solver
def test_two_gpu_sweep_matches_reference(name):
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    from test_gpu_dropin import sweep_bounds
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "scripts", "run_dropin_multigpu.py"), name]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1800, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-3000:]
    res = json.loads(lines[-1])
    assert res["rc"] == [0, 0], res
    assert res["ranks_identical"], res
    bounds = sweep_bounds(name, res["sweeps"])
    assert res["max_abs_dE_vs_reference"] is not None and res["max_abs_dE_vs_reference"] <= max(bounds), res
