"""2-rank NCCL run of the product's multi-GPU path on real GPUs: batch rows sharded over the ranks, context
broadcast from rank 0, images gathered — every rank's rows must equal, BIT FOR BIT, the rows a single process produces for
the same global row indices at the same per-GPU batch.  Skipped on a 1-GPU lease (the CPU twin is tests/test_dist_cpu.py)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_run_reproduces_single_rank_rows(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import dist_gpu_worker as W
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_gpu_worker.py"), str(tmp_path)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    r = [torch.load(tmp_path / f"rank{i}.pt") for i in range(2)]
    assert r[0]["rows"] == (0, W.BS) and r[1]["rows"] == (W.BS, 2 * W.BS)
    from test_parity_gpu import build_net
    net, _ = build_net(mini=True)
    dev = torch.device("cuda", 0)
    ctx = W.make_context(dev, fill=True)
    for i in range(2):
        x, img = W.sample_rows(net, r[i]["rows"], ctx, dev)
        assert torch.equal(x, r[i]["x"]), f"rank {i}: latents differ from the single-process rows"
        assert torch.equal(img, r[i]["img"]), f"rank {i}: images differ from the single-process rows"
    assert r[1]["gathered"] is None and torch.equal(r[0]["gathered"], torch.cat([r[0]["img"], r[1]["img"]]))
