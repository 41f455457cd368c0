from .combine_jastrow import CombineJastrow  # noqa: F401
from .elec_elec import JastrowFactorElectronElectron  # noqa: F401
from .elec_nuclei import JastrowFactorElectronNuclei  # noqa: F401
from .elec_elec_nuclei import JastrowFactorElectronElectronNuclei  # noqa: F401
