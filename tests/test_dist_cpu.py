"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: batch sharding, per-global-row latents,
context broadcast and image gather (vdb200/parallel.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "versatile-diffusion_b200"))
    from vdb200 import parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rows = parallel.shard_rows(7)                                    # ragged: 4 + 3
    xT = parallel.seeded_latents(rows, (4, 8, 8), seed=123)
    ctx = torch.arange(12.).view(1, 3, 4) if rank == 0 else torch.zeros(1, 3, 4)
    parallel.broadcast_context([ctx])
    imgs = parallel.gather_images(xT * 2.0)
    torch.save({"rows": rows, "xT": xT, "ctx": ctx, "imgs": imgs}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank(tmp_path):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "versatile-diffusion_b200"))
    from vdb200 import parallel
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    assert r0["rows"] == (0, 4) and r1["rows"] == (4, 7)
    full = parallel.seeded_latents((0, 7), (4, 8, 8), seed=123)      # what a 1-rank run draws
    assert torch.equal(torch.cat([r0["xT"], r1["xT"]]), full)
    assert torch.equal(r0["ctx"], r1["ctx"]) and r1["ctx"].sum() == 66.0
    assert torch.equal(r0["imgs"], full * 2.0) and r1["imgs"] is None
