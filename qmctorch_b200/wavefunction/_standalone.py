"""Plan handles for operators used on their own (outside a SlaterJastrow), e.g. a Jastrow
factor evaluated directly on a duck-typed ``mol`` as in the reference's unit tests
(tests/wavefunction/jastrows/elec_elec/test_pade_jastrow.py:19-32)."""
from types import SimpleNamespace

import numpy as np

from ._plan import PlanHandle


def _placeholder_molecule(mol):
    """Smallest valid basis so that the shared tables can be built when ``mol`` has none."""
    coords = np.asarray(getattr(mol, "atom_coords", [[0.0, 0.0, 0.0]]), dtype=np.float64).reshape(-1, 3)
    natom = coords.shape[0]
    n = max(mol.nup, mol.ndown, 1)
    b = SimpleNamespace(
        nao=n, nmo=n, nshells=[n] + [0] * (natom - 1), nao_per_atom=[n] + [0] * (natom - 1),
        index_ctr=list(range(n)), nctr_per_ao=np.ones(n, dtype=int), bas_coeffs=np.ones(n),
        bas_exp=1.0 + np.arange(n, dtype=np.float64), bas_kr=np.zeros(n), bas_kx=np.zeros(n, dtype=int),
        bas_ky=np.zeros(n, dtype=int), bas_kz=np.zeros(n, dtype=int), radial_type="gto_pure",
        harmonics_type="cart", mos=np.eye(n), atom_coords_internal=coords.tolist())
    return SimpleNamespace(nelec=mol.nup + mol.ndown, nup=mol.nup, ndown=mol.ndown, basis=b,
                           atomic_number=list(getattr(mol, "atomic_number", [1] * natom)))


def standalone_handle(mol, owner, jee=None, jen=None, jeen=None):
    from .orbitals.atomic_orbitals import AtomicOrbitals
    if not hasattr(mol, "basis") or not hasattr(mol.basis, "bas_exp"):
        mol = _placeholder_molecule(mol)
    ao = AtomicOrbitals(mol, cuda=True)
    owner.__dict__["_standalone_ao"] = ao      # keep alive without registering as a sub-module
    return PlanHandle(ao, jastrow_ee=jee, jastrow_en=jen, jastrow_een=jeen, nup=mol.nup, ndown=mol.ndown)
