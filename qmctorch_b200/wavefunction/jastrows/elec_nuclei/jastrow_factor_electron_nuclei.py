"""Electron-nucleus Jastrow factor operator
(qmctorch/wavefunction/jastrows/elec_nuclei/jastrow_factor_electron_nuclei.py:9-161)."""
import numpy as np
import torch
from torch import nn

from .._base import jastrow_forward
from .kernels import PadeJastrowKernel


class JastrowFactorElectronNuclei(nn.Module):
    def __init__(self, mol, jastrow_kernel, kernel_kwargs={}, cuda=False):
        super().__init__()
        self.nup, self.ndown = mol.nup, mol.ndown
        self.nelec = mol.nup + mol.ndown
        self.cuda = cuda
        self.device = torch.device("cuda", torch.cuda.current_device()) if cuda else torch.device("cpu")
        atomic_pos = torch.as_tensor(np.asarray(mol.atom_coords), dtype=torch.float64)
        self.atoms = atomic_pos.to(self.device)
        self.natoms = self.atoms.shape[0]
        self.ndim = 3
        self.jastrow_kernel = jastrow_kernel(mol.nup, mol.ndown, atomic_pos, cuda, **kernel_kwargs)
        if not isinstance(self.jastrow_kernel, PadeJastrowKernel):
            raise NotImplementedError("only the analytic PadeJastrowKernel is fused into the CUDA path")
        self.requires_autograd = self.jastrow_kernel.requires_autograd
        self._handle = None
        self._mol = mol

    def __repr__(self):
        return "en -> " + self.jastrow_kernel.__class__.__name__

    def _own_handle(self):
        if self._handle is None:
            from ..._standalone import standalone_handle
            self._handle = standalone_handle(self._mol, self, jen=self)
        return self._handle

    def forward(self, pos, derivative=0, sum_grad=True):
        return jastrow_forward(self, self._own_handle(), 2, pos, derivative, sum_grad)
