"""DiagonalGaussianDistribution (reference lib/model_zoo/distributions.py:24-62) — API-surface object returned by
AutoencoderKL.encode(out_posterior=True).  The sampling hot path does not build it: encode() samples through
the fused vdb200 gaussian_sample kernel."""
import torch


class DiagonalGaussianDistribution(object):
    def __init__(self, parameters, deterministic=False):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)
        if self.deterministic:
            self.var = self.std = torch.zeros_like(self.mean).to(device=self.parameters.device)

    def sample(self):
        # the reference draws on the CPU generator and then moves (distributions.py:36)
        return self.mean + self.std * torch.randn(self.mean.shape).to(device=self.parameters.device)

    def mode(self):
        return self.mean
