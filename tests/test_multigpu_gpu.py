"""Multi-GPU correctness inside pytest: the sharded BA (points), GP (points) and RA (edges) must
reproduce the single-GPU solves -- same cost after every one of a fixed number of LM iterations, same natural
iteration count, same converged solution.  The checks themselves are the torchrun scripts tests/multigpu_*_check.py
(one process per GPU, NCCL); this module launches them on 2 GPUs (and on all visible GPUs when there are more) and
is skipped on a single-GPU box -- two NCCL ranks cannot share one device."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _run(script, nproc, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", script)]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):          # keep the full transcript of the ranks next to the other GPU-box artefacts
        with open(os.path.join(out_dir, f"multigpu_{script}_{nproc}.log"), "w") as f:
            f.write(r.stdout)
    # the ranks' own lines first (the launcher's traceback at the end says nothing about the cause)
    own = [ln for ln in r.stdout.splitlines() if not ln.startswith(("E  ", "  File", "    ")) and "torch/distributed" not in ln]
    assert r.returncode == 0, "\n".join(own[-60:])
    return r.stdout


@pytest.mark.parametrize("nproc", [2, 8])
def test_sharded_ba_matches_single_gpu(nproc):
    n = _ngpu()
    if n < nproc:
        pytest.skip(f"needs {nproc} GPUs on one box, {n} visible (NCCL ranks cannot share a device)")
    out = _run("multigpu_ba_check.py", nproc, 29511 + nproc)
    assert "multi-GPU parity OK" in out, out[-2000:]


@pytest.mark.parametrize("nproc", [2, 8])
def test_sharded_gp_and_ra_match_single_gpu(nproc):
    n = _ngpu()
    if n < nproc:
        pytest.skip(f"needs {nproc} GPUs on one box, {n} visible (NCCL ranks cannot share a device)")
    out = _run("multigpu_gp_ra_check.py", nproc, 29531 + nproc)
    assert "multi-GPU GP/RA parity OK" in out, out[-2000:]
