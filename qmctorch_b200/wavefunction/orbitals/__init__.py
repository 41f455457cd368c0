from .atomic_orbitals import AtomicOrbitals  # noqa: F401
from .molecular_orbitals import MolecularOrbitals  # noqa: F401
