"""Multi-GPU parity inside `pytest -m gpu`: spawns tools/ddp_check.py under torchrun for
world 2 / 4 / 8 when that many GPUs are visible (skipped otherwise): ct_allreduce_bucket / ct_broadcast bit-identical to
NCCL (integer-valued data) in every available mode incl. NVLS multimem, bit-identical across ranks; the DDP wrapper's
gradients vs the mean of per-rank gradients, vs torch DDP over NCCL around the oracle (the reference's recipe,
examples/ft_bloom_DDP.py:99,145-150), tied-table dense + sparse split, GPT with segment_ids, and the CUDA-graph replay
of the whole DDP step."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


@pytest.mark.parametrize("world", [2, 4, 8])
def test_ddp_and_collectives_parity(world, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs, %d visible" % (world, torch.cuda.device_count()))
    out = tmp_path / "ddp_check.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tools", "ddp_check.py"), "--quick", "--out", str(out)]
    env = dict(os.environ, CT_COMM_TIMEOUT_S="120")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    res = json.load(open(out))
    assert res["world"] == world and res["failures_all_ranks"] == []
    assert res["ddp"]["p2p_vs_mean"] <= 4e-3 and res["ddp"]["graph_vs_eager"] <= 4e-3
    try:  # keep the numbers next to the other evidence when the directory is writable
        dst = os.path.join(ROOT, "gpurun_out", "r02_ddp_check_w%d.json" % world)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        json.dump(res, open(dst, "w"), indent=1)
    except OSError:
        pass
