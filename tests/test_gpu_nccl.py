"""GPU, multi-rank: BASELINE configs[3] over NCCL inside the driver's own `pytest -m gpu` run.

Spawns `python -m torch.distributed.run` over min(device_count, 8) ranks on tools/nccl_parity.py: a 32 768-set batch
sharded by the parallel_chunks rule, NCCL all-gather of the 576-byte partials, one final exponentiation — valid batch
true on every rank, corrupted batch false with the GT a single context computes (and BLST's on a window), and the
sharded G1 MSM equal to the single-context MSM and to blst_p1s_mult_pippenger.  Skipped below two GPUs (the one-GPU
equivalents are tests/test_gpu_golden_and_shares.py::test_rank_shares_on_one_device and tests/test_gpu_abi_multi.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ndev():
    import nim_blscurve_b200 as bg
    return bg.lib().blsgpu_device_count()


@pytest.mark.parametrize("n", [32768])
def test_nccl_batch_verify_and_msm(n):
    world = min(_ndev(), 8)
    if world < 2:
        pytest.skip("needs at least two GPUs (NCCL ranks)")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "nccl_parity.py"), str(n)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert f"nccl_parity OK: {n} sets over {world} ranks" in r.stdout
    assert "== single context == BLST" in r.stdout
