"""Data parallelism through the PRODUCT path on a GPU: two ranks drive model.sggan.train_step (the world_size > 1 branch:
phase-split step, all-reduce of the two flat gradient buffers, 1/world folded into Adam) and must land where ONE rank
lands on the concatenated batch (instance norm is per-sample, losses are batch means: SURVEY 8(e)).

Both ranks share cuda:0 and talk over gloo (which moves CUDA tensors through the host), so the test runs on the
driver's single-GPU box; `tests/gpu/dp_worker.py --backend nccl` under torchrun on 2 GPUs exercises the same code
over NCCL (log in profiles/)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(900)
@pytest.mark.parametrize("buckets", ["0", "1"])
def test_two_ranks_match_one_rank_on_the_concatenated_batch(buckets):
    """buckets "0": one all-reduce of G's gradients after the backward (default); "1": the backward in two parts with the
    upper bucket reduced underneath the second part (SGGAN_DP_BUCKETS=1, sggan_step_backward_g_part)."""
    port = 29600 + (os.getpid() % 300) + int(buckets)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), SGGAN_DP_BUCKETS=buckets)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "gpu", "dp_worker.py"), "--backend", "gloo", "--one-gpu"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=850, env=env)
    assert out.returncode == 0 and "DP-OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
