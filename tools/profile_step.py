"""One eager DDIM step (+ optionally the VAE decode) of the C2 workload inside a cudaProfiler range, for ncu:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/profile_step.py
"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "versatile-diffusion_b200"))
import torch
import bench
from lib.model_zoo.ddim import DDIMSampler

dev = torch.device("cuda", 0)
net = bench.build_net(dev)
S = DDIMSampler(net, use_cuda_graph=False)
g = torch.Generator().manual_seed(0)
xT = torch.randn(4, 4, 64, 64, generator=g).to(dev)
c = (torch.randn(4, 77, 768, generator=g) * 0.5).to(dev); u = (torch.randn(4, 77, 768, generator=g) * 0.5).to(dev)
def run(steps):
    with torch.no_grad():
        x, _ = S.sample(steps=steps, shape=[4, 4, 64, 64], x_info={"type": "image", "xt": xT},
                        c_info={"type": "text", "conditioning": c, "unconditional_conditioning": u, "unconditional_guidance_scale": 7.5},
                        verbose=False, eta=0.)
        return x
x = run(2)
if "--decode" in sys.argv:
    net.vae_decode(x, "image")
torch.cuda.synchronize()
torch.cuda.profiler.start()
x = run(1)
if "--decode" in sys.argv:
    net.vae_decode(x, "image")
torch.cuda.synchronize()
torch.cuda.profiler.stop()
