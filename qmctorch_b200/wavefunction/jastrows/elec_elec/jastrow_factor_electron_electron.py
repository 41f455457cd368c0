"""Electron-electron Jastrow factor operator
(qmctorch/wavefunction/jastrows/elec_elec/jastrow_factor_electron_electron.py:14-260)."""
import torch
from torch import nn

from .._base import jastrow_forward
from .kernels import PadeJastrowKernel


class JastrowFactorElectronElectron(nn.Module):
    def __init__(self, mol, jastrow_kernel, kernel_kwargs={}, scale=False, scale_factor=0.6, cuda=False):
        super().__init__()
        self.nup, self.ndown = mol.nup, mol.ndown
        self.nelec = mol.nup + mol.ndown
        self.ndim = 3
        self.cuda = cuda
        self.device = torch.device("cuda", torch.cuda.current_device()) if cuda else torch.device("cpu")
        if scale:
            raise NotImplementedError("scaled distances (distance/scaling.py) are not on the CUDA path")
        self.jastrow_kernel = jastrow_kernel(mol.nup, mol.ndown, cuda, **kernel_kwargs)
        if not isinstance(self.jastrow_kernel, PadeJastrowKernel):
            raise NotImplementedError(
                "only the analytic PadeJastrowKernel is fused into the CUDA path; kernels that need "
                "autograd (fully connected, polynomial Pade) are outside BASELINE north_star")
        self.requires_autograd = self.jastrow_kernel.requires_autograd
        self._handle = None
        self._mol = mol

    def __repr__(self):
        return "ee -> " + self.jastrow_kernel.__class__.__name__

    def _own_handle(self):
        if self._handle is None:
            from ..._standalone import standalone_handle
            self._handle = standalone_handle(self._mol, self, jee=self)
        return self._handle

    def forward(self, pos, derivative=0, sum_grad=True):
        return jastrow_forward(self, self._own_handle(), 1, pos, derivative, sum_grad)
