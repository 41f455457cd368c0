"""``mol`` data contract (stands in for qmctorch/scf; the SCF front end itself is out of scope)."""
from ..molecules import Molecule, fixture_molecule  # noqa: F401
