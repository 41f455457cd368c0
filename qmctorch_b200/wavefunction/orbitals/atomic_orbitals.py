"""AtomicOrbitals operator - same constructor, attributes and forward contract as
qmctorch/wavefunction/orbitals/atomic_orbitals.py:15-218; the arithmetic runs in
``qmcb_ao`` (csrc/operators.cu), one thread per (walker, electron) with the regrouped
basis in shared memory."""
import numpy as np
import torch
from torch import nn

from ... import _lib
from .._plan import PlanHandle, as_walkers
from .norm_orbital import atomic_orbital_norm


class AtomicOrbitals(nn.Module):
    def __init__(self, mol, cuda=False):
        super().__init__()
        dtype = torch.float64
        basis = mol.basis
        self.nelec = mol.nelec
        self.nup, self.ndown = mol.nup, mol.ndown
        self.norb = basis.nao
        self.ndim = 3
        self.atom_coords = nn.Parameter(torch.as_tensor(np.asarray(basis.atom_coords_internal), dtype=dtype))
        self.atom_coords.requires_grad = True
        self.natoms = len(self.atom_coords)
        self.atomic_number = mol.atomic_number
        self.nshells = torch.as_tensor(np.asarray(basis.nshells))
        self.nao_per_atom = torch.as_tensor(np.asarray(basis.nao_per_atom))
        self.nbas = int(self.nshells.sum())
        self.index_ctr = torch.as_tensor(np.asarray(basis.index_ctr))
        self.nctr_per_ao = torch.as_tensor(np.asarray(basis.nctr_per_ao))
        self.contract = not len(torch.unique(self.index_ctr)) == len(self.index_ctr)
        self.bas_coeffs = torch.as_tensor(np.asarray(basis.bas_coeffs), dtype=dtype)
        self.bas_exp = nn.Parameter(torch.as_tensor(np.asarray(basis.bas_exp), dtype=dtype))
        self.bas_exp.requires_grad = True
        self.harmonics_type = basis.harmonics_type
        if basis.harmonics_type != "cart":
            raise NotImplementedError(
                "harmonics_type='sph' (spherical_harmonics.py:202-702) is not on the CUDA path")
        self.bas_n = torch.as_tensor(np.asarray(basis.bas_kr), dtype=dtype)
        self.radial_type = basis.radial_type
        if self.radial_type not in _lib.RADIAL:
            raise ValueError("unknown radial_type %r" % self.radial_type)
        with torch.no_grad():
            self.norm_cst = torch.as_tensor(atomic_orbital_norm(basis), dtype=dtype)
        # host-side integer tables handed to the plan
        self.bas_atom_np = np.repeat(np.arange(self.natoms), np.asarray(basis.nshells)).astype(np.int32)
        self.bas_kx_np = np.asarray(basis.bas_kx).astype(np.int32)
        self.bas_ky_np = np.asarray(basis.bas_ky).astype(np.int32)
        self.bas_kz_np = np.asarray(basis.bas_kz).astype(np.int32)
        self.bas_kr_np = np.asarray(basis.bas_kr).astype(np.int32)
        self.index_ctr_np = np.asarray(basis.index_ctr).astype(np.int32)
        self.backflow_trans = None
        self.cuda = cuda
        self.device = torch.device("cpu")
        self._handle = PlanHandle(self)
        if self.cuda:
            self._to_device()

    def __repr__(self):
        return self.__class__.__name__ + "(%s, %s, %d -> (%d,%d) )" % (
            self.radial_type, self.harmonics_type, self.nelec * self.ndim, self.nelec, self.norb)

    def _to_device(self):
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.to(self.device)
        for at in ["bas_n", "bas_coeffs", "nshells", "norm_cst", "index_ctr", "nctr_per_ao",
                   "nao_per_atom"]:
            self.__dict__[at] = self.__dict__[at].to(self.device)

    def forward(self, pos, derivative=[0], sum_grad=True, sum_hess=True, one_elec=False):
        """pos [W, 3*nelec] -> ao [W,ne,nao]; derivative=1 -> summed or [W,ne,nao,3];
        derivative=2 -> Laplacian; [0,1,2] -> (ao, dao[...,3], d2ao)  (atomic_orbitals.py:131-218)."""
        if not isinstance(derivative, list):
            derivative = [derivative]
        if not sum_grad:
            assert 1 in derivative
        if not sum_hess:
            raise NotImplementedError(
                "individual second derivatives (sum_hess=False, atomic_orbitals.py:453-516) are only "
                "used by backflow and are not on the CUDA path")
        if derivative not in ([0], [1], [2], [0, 1, 2]):
            if derivative == [3]:
                raise NotImplementedError("mixed second derivatives (backflow only) are not on the CUDA path")
            raise ValueError("derivative must be 0, 1, 2, 3 or [0, 1, 2, 3], got ", derivative)
        ne = 1 if one_elec else self.nelec
        dev = self.atom_coords.device
        x = as_walkers(pos, 3 * ne, dev)
        W = x.shape[0]
        L = _lib.lib()
        plan = self._handle.plan()
        ao = torch.empty(W, ne, self.norb, dtype=torch.float64, device=dev)
        if derivative == [0]:
            _lib.check(L.qmcb_ao(plan, _lib.ptr(x), W, int(one_elec), _lib.ptr(ao), None, None,
                                 _lib.stream_ptr(dev)), "qmcb_ao")
            return ao
        dao = torch.empty(W, ne, self.norb, 3, dtype=torch.float64, device=dev)
        d2ao = torch.empty(W, ne, self.norb, dtype=torch.float64, device=dev)
        _lib.check(L.qmcb_ao(plan, _lib.ptr(x), W, int(one_elec), _lib.ptr(ao), _lib.ptr(dao),
                             _lib.ptr(d2ao), _lib.stream_ptr(dev)), "qmcb_ao")
        if derivative == [1]:
            return dao.sum(-1) if sum_grad else dao
        if derivative == [2]:
            return d2ao
        return ao, dao, d2ao

    def update(self, ao, pos, idelec):
        """atomic_orbitals.py:671-695."""
        ao_new = ao.clone()
        ids, ide = idelec * 3, (idelec + 1) * 3
        ao_new[:, idelec, :] = self.forward(pos[:, ids:ide], one_elec=True).squeeze(1)
        return ao_new
