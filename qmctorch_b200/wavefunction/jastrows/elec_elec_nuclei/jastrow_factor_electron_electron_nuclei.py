"""Three-body electron-electron-nucleus Jastrow factor operator
(qmctorch/wavefunction/jastrows/elec_elec_nuclei/jastrow_factor_electron_electron_nuclei.py:16-439)."""
import numpy as np
import torch
from torch import nn

from .._base import jastrow_forward
from .kernels import BoysHandyJastrowKernel


class JastrowFactorElectronElectronNuclei(nn.Module):
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
        if not isinstance(self.jastrow_kernel, BoysHandyJastrowKernel):
            raise NotImplementedError("only the BoysHandyJastrowKernel is fused into the CUDA path")
        self.requires_autograd = self.jastrow_kernel.requires_autograd
        self.auto_second_derivative = False     # analytic Laplacian in the kernel
        self._handle = None
        self._mol = mol
        if cuda:
            self.to(self.device)

    def __repr__(self):
        return "een -> " + self.jastrow_kernel.__class__.__name__

    def _own_handle(self):
        if self._handle is None:
            from ..._standalone import standalone_handle
            self._handle = standalone_handle(self._mol, self, jeen=self)
        return self._handle

    def forward(self, pos, derivative=0, sum_grad=True):
        return jastrow_forward(self, self._own_handle(), 3, pos, derivative, sum_grad)
