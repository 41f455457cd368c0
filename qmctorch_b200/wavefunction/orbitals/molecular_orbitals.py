"""MolecularOrbitals operator (qmctorch/wavefunction/orbitals/molecular_orbitals.py:8-95):
out = ao @ (mo_scf * mo_modifier), evaluated by ``qmcb_mo``.  ``mo_modifier`` is the
trainable parameter; inside the fused kernel only the columns some configuration
occupies are ever formed."""
import numpy as np
import torch
from torch import nn

from ... import _lib


class MolecularOrbitals(nn.Module):
    def __init__(self, mol, include_all_mo, highest_occ_mo, mix_mo, orthogonalize_mo, cuda):
        super().__init__()
        if mix_mo:
            # the reference itself crashes here (molecular_orbitals.py:49 assigns to None.weight)
            raise NotImplementedError("mix_mo=True is broken in the reference and not provided")
        if orthogonalize_mo:
            raise Warning("orthogonalize_mo=True has no effect as mix_mo=False")
        self.mol = mol
        self.mix_mo = mix_mo
        self.orthogonalize_mo = orthogonalize_mo
        self.cuda = cuda
        self.device = torch.device("cpu")
        self.include_all_mo = include_all_mo
        self.highest_occ_mo = highest_occ_mo
        self.nmo_opt = mol.basis.nmo if include_all_mo else highest_occ_mo
        mo = torch.as_tensor(np.asarray(mol.basis.mos), dtype=torch.float64)
        if not include_all_mo:
            mo = mo[:, :highest_occ_mo]
        self.mo_scf = mo.contiguous().requires_grad_(False)
        self.mo_modifier = nn.Parameter(torch.ones_like(self.mo_scf))
        self.mo_mixer = None
        self._handle = None
        if cuda:
            self.device = torch.device("cuda", torch.cuda.current_device())
            self.mo_scf = self.mo_scf.to(self.device)
            self.to(self.device)

    def get_mo_coeffs(self):
        return self.mo_scf

    def forward(self, ao):
        """ao [..., nao] -> [..., nmo]  (molecular_orbitals.py:79-95)."""
        if self._handle is None:
            raise RuntimeError("MolecularOrbitals must be attached to a SlaterJastrow")
        dev = self.mo_modifier.device
        x = ao.detach().to(device=dev, dtype=torch.float64).contiguous()
        rows = x.numel() // x.shape[-1]
        out = torch.empty(*x.shape[:-1], self.mo_scf.shape[1], dtype=torch.float64, device=dev)
        L = _lib.lib()
        _lib.check(L.qmcb_mo(self._handle.plan(), _lib.ptr(x), rows, _lib.ptr(out), _lib.stream_ptr(dev)),
                   "qmcb_mo")
        return out
