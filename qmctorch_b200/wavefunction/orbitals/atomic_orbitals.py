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


# Real spherical harmonics up to l = 2 as the reference defines them (spherical_harmonics.py:352-702,
# including its truncated literal for Y00): Y_lm = sum_t c_t x^a y^b z^c / r^l.  With the radial part
# r^n exp(-alpha r) or r^n exp(-alpha r^2) (radial_functions.py:6-238) the primitive is a sum of CARTESIAN
# monomials with radial power n - l, which is what the kernels evaluate:  (l, m) -> [(c, a, b, c)]
_C1, _C20, _C22, _C2M = 0.4886025119029199, 0.31539156525252005, 0.5462742152960396, 1.0925484305920792
_SPH = {
    (0, 0): [(0.2820948, 0, 0, 0)],
    (1, -1): [(_C1, 0, 1, 0)], (1, 0): [(_C1, 0, 0, 1)], (1, 1): [(_C1, 1, 0, 0)],
    (2, -2): [(_C2M, 1, 1, 0)], (2, -1): [(_C2M, 0, 1, 1)], (2, 1): [(_C2M, 1, 0, 1)],
    (2, 2): [(_C22, 2, 0, 0), (-_C22, 0, 2, 0)],
    (2, 0): [(-_C20, 2, 0, 0), (-_C20, 0, 2, 0), (2.0 * _C20, 0, 0, 2)],
}


def _expand_spherical(basis):
    """Flat primitives of a spherical-harmonics basis -> (index of the original primitive, monomial
    coefficient, kx, ky, kz, radial power) of the equivalent cartesian primitives."""
    n = np.asarray(basis.bas_n).astype(int)
    lq = np.asarray(basis.bas_l).astype(int)
    mq = np.asarray(basis.bas_m).astype(int)
    if basis.radial_type.endswith("pure") and np.any(lq > 0):
        raise NotImplementedError(
            "spherical harmonics with l > 0 need a radial power n - l: use radial_type 'sto' or 'gto' "
            "(the *_pure radial functions carry no r^n factor)")
    idx, scale, kx, ky, kz, kr = [], [], [], [], [], []
    for i, (ni, li, mi) in enumerate(zip(n, lq, mq)):
        if (li, mi) not in _SPH:
            raise NotImplementedError("spherical harmonics are implemented up to l = 2 (as in the reference); got "
                                      "l = %d, m = %d" % (li, mi))
        if ni < li:
            raise NotImplementedError("radial power n = %d below l = %d: r^(n-l) Y_lm would be singular" % (ni, li))
        for c, a, b, cz in _SPH[(li, mi)]:
            idx.append(i); scale.append(c); kx.append(a); ky.append(b); kz.append(cz); kr.append(ni - li)
    i32 = lambda v: np.asarray(v, dtype=np.int32)
    return np.asarray(idx, dtype=np.int64), np.asarray(scale, dtype=np.float64), i32(kx), i32(ky), i32(kz), i32(kr)


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
        self.nbas = int(self.nshells.sum())          # the reference's flat primitives (len(bas_exp))
        self.index_ctr = torch.as_tensor(np.asarray(basis.index_ctr))
        self.nctr_per_ao = torch.as_tensor(np.asarray(basis.nctr_per_ao))
        self.contract = not len(torch.unique(self.index_ctr)) == len(self.index_ctr)
        self.bas_coeffs = torch.as_tensor(np.asarray(basis.bas_coeffs), dtype=dtype)
        self.bas_exp = nn.Parameter(torch.as_tensor(np.asarray(basis.bas_exp), dtype=dtype))
        self.bas_exp.requires_grad = True
        self.harmonics_type = basis.harmonics_type
        self.radial_type = basis.radial_type
        if self.radial_type not in _lib.RADIAL:
            raise ValueError("unknown radial_type %r" % self.radial_type)
        with torch.no_grad():
            self.norm_cst = torch.as_tensor(atomic_orbital_norm(basis), dtype=dtype)
        bas_atom = np.repeat(np.arange(self.natoms), np.asarray(basis.nshells)).astype(np.int32)
        index_ctr = np.asarray(basis.index_ctr).astype(np.int32)
        # primitives handed to the plan: one per (primitive, cartesian monomial).  expand_index maps them back
        # to the reference's flat primitives, expand_scale is the monomial's coefficient (folded into the norm)
        self.expand_index = None
        self.expand_scale = None
        if basis.harmonics_type == "cart":
            self.bas_n = torch.as_tensor(np.asarray(basis.bas_kr), dtype=dtype)
            kx, ky, kz, kr = (np.asarray(getattr(basis, k)).astype(np.int32) for k in ("bas_kx", "bas_ky", "bas_kz", "bas_kr"))
        elif basis.harmonics_type == "sph":
            self.bas_n = torch.as_tensor(np.asarray(basis.bas_n), dtype=dtype)
            idx, scale, kx, ky, kz, kr = _expand_spherical(basis)
            self.expand_index, self.expand_scale = idx, scale
            bas_atom, index_ctr = bas_atom[idx], index_ctr[idx]
        else:
            raise ValueError("harmonics_type should be 'cart' or 'sph'")
        # host-side integer tables handed to the plan
        self.bas_atom_np = bas_atom
        self.bas_kx_np, self.bas_ky_np, self.bas_kz_np, self.bas_kr_np = kx, ky, kz, kr
        self.index_ctr_np = index_ctr
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
