"""Worker of tests/test_dist_gpu.py (launched by torch.distributed.run, one process per GPU): samples this rank's rows of a
global batch through the product path (vdb200.parallel plumbing + DDIMSampler + vae_decode) and saves them."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "versatile-diffusion_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

BS, SEED, STEPS = 2, 77, 4


def sample_rows(net, rows, ctx, device):
    from lib.model_zoo.ddim import DDIMSampler
    from vdb200 import parallel
    xT = parallel.seeded_latents(rows, (4, 16, 16), seed=SEED).to(device)
    c, u = ctx
    n = rows[1] - rows[0]
    with torch.no_grad():
        x, _ = DDIMSampler(net).sample(steps=STEPS, shape=[n, 4, 16, 16], x_info={"type": "image", "xt": xT},
                                       c_info={"type": "text", "conditioning": c.repeat(n, 1, 1), "unconditional_conditioning": u.repeat(n, 1, 1),
                                               "unconditional_guidance_scale": 7.5}, verbose=False, eta=0.)
        img = net.vae_decode(x, "image")
    return x.float().cpu(), img.float().cpu()


def make_context(device, fill):
    g = torch.Generator().manual_seed(9)
    c, u = torch.randn(1, 77, 768, generator=g) * 0.5, torch.randn(1, 77, 768, generator=g) * 0.5
    if not fill:
        c, u = torch.zeros_like(c), torch.zeros_like(u)
    return c.to(device), u.to(device)


def main(out_dir):
    from test_parity_gpu import build_net
    from vdb200 import parallel
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    net, _ = build_net(mini=True)
    ctx = make_context(device, fill=(rank == 0))           # only rank 0 "encodes"; the others receive the broadcast
    parallel.broadcast_context(list(ctx))
    rows = parallel.shard_rows(BS * world)
    x, img = sample_rows(net, rows, ctx, device)
    gathered = parallel.gather_images(img.to(device))
    torch.save({"rows": rows, "x": x, "img": img, "gathered": None if gathered is None else gathered.cpu()},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
