"""Electron-nucleus Jastrow kernels (elec_nuclei/kernels/pade_jastrow_kernel.py:8-116)."""
import torch
from torch import nn


class JastrowKernelElectronNucleiBase(nn.Module):
    def __init__(self, nup, ndown, atomic_pos, cuda, **kwargs):
        super().__init__()
        self.nup, self.ndown = nup, ndown
        self.cuda = cuda
        self.device = torch.device("cuda", torch.cuda.current_device()) if cuda else torch.device("cpu")
        self.atoms = atomic_pos
        self.natoms = atomic_pos.shape[0]
        self.ndim = 3
        self.requires_autograd = True

    def forward(self, r):
        raise NotImplementedError()


class PadeJastrowKernel(JastrowKernelElectronNucleiBase):
    """r / (1 + w r) with static weight 1."""

    def __init__(self, nup, ndown, atomic_pos, cuda, w=1.0):
        super().__init__(nup, ndown, atomic_pos, cuda)
        self.weight = nn.Parameter(torch.as_tensor([w], dtype=torch.float64), requires_grad=True)
        self.static_weight = torch.as_tensor([1.0], dtype=torch.float64).to(self.device)
        self.requires_autograd = True
