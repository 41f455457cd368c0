"""Electron-electron Jastrow kernels.  Only the analytic Pade kernel is fused into the CUDA
path; the plug-in base class keeps the reference contract
(jastrow_kernel_electron_electron_base.py:6-107)."""
import torch
from torch import nn


class JastrowKernelElectronElectronBase(nn.Module):
    def __init__(self, nup, ndown, cuda, **kwargs):
        super().__init__()
        self.nup, self.ndown = nup, ndown
        self.cuda = cuda
        self.device = torch.device("cuda", torch.cuda.current_device()) if cuda else torch.device("cpu")
        self.requires_autograd = True

    def forward(self, r):
        raise NotImplementedError()


class PadeJastrowKernel(JastrowKernelElectronElectronBase):
    """w0 r / (1 + w r), w0 = 0.25 same spin block / 0.5 otherwise as coded in
    elec_elec/kernels/pade_jastrow_kernel.py:34-66; ``weight`` is the trainable w."""

    def __init__(self, nup, ndown, cuda, w=1.0):
        super().__init__(nup, ndown, cuda)
        self.weight = nn.Parameter(torch.as_tensor([w], dtype=torch.float64), requires_grad=True)
        self.requires_autograd = False
        bup = torch.cat((0.25 * torch.ones(nup, nup), 0.5 * torch.ones(nup, ndown)), dim=1)
        bdown = torch.cat((0.5 * torch.ones(ndown, nup), 0.25 * torch.ones(ndown, ndown)), dim=1)
        sw = torch.cat((bup, bdown), dim=0)
        self.static_weight = sw[torch.triu(torch.ones_like(sw), diagonal=1).bool()].to(self.device)
