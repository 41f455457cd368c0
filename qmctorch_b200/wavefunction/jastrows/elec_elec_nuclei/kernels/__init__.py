"""Electron-electron-nucleus Jastrow kernels.  The Boys-Handy kernel
(elec_elec_nuclei/kernels/boys_handy_jastrow_kernel.py:8-93) is fused into the CUDA path with
hand-derived gradient and Laplacian (the reference obtains both by autograd)."""
import torch
from torch import nn


class JastrowKernelElectronElectronNucleiBase(nn.Module):
    def __init__(self, nup, ndown, atomic_pos, cuda, **kwargs):
        super().__init__()
        self.nup, self.ndown = nup, ndown
        self.cuda = cuda
        self.nelec = nup + ndown
        self.atoms = atomic_pos
        self.natoms = atomic_pos.shape[0]
        self.ndim = 3
        self.device = torch.device("cuda", torch.cuda.current_device()) if cuda else torch.device("cpu")
        self.requires_autograd = True

    def forward(self, x):
        raise NotImplementedError()


class BoysHandyJastrowKernel(JastrowKernelElectronElectronNucleiBase):
    """K(r_iA, r_jA, r_ij) = sum_mu c_mu f_mu(r_iA) f_mu(r_jA) g_mu(r_ij); same parameter names,
    shapes and initial values as the reference: weight_num [1,2,nterm] = a0, weight_denom
    [1,2,nterm] = b0, fc = nn.Linear(nterm, 1, bias=False)."""

    MAXTERM = 8

    def __init__(self, nup, ndown, atomic_pos, cuda, a0=1e-3, b0=1.0, exp0=1.0, nterm=5):
        super().__init__(nup, ndown, atomic_pos, cuda)
        if exp0 != 1.0:
            raise NotImplementedError("the CUDA path implements the Boys-Handy kernel with exponents 1 (default)")
        if nterm > self.MAXTERM:
            raise NotImplementedError("at most %d Boys-Handy terms" % self.MAXTERM)
        self.nterm = nterm
        self.fc = nn.Linear(nterm, 1, bias=False).to(torch.float64)
        self.weight_num = nn.Parameter(a0 * torch.ones(1, 2, nterm, dtype=torch.float64), requires_grad=True)
        self.weight_denom = nn.Parameter(b0 * torch.ones(1, 2, nterm, dtype=torch.float64), requires_grad=True)
        self.exp = exp0 * torch.ones(2, nterm, dtype=torch.float64)
