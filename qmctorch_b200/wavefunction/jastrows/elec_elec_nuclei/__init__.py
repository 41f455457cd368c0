from .jastrow_factor_electron_electron_nuclei import JastrowFactorElectronElectronNuclei  # noqa: F401
from .jastrow_factor_electron_electron_nuclei import JastrowFactorElectronElectronNuclei as JastrowFactor  # noqa: F401
from .kernels import BoysHandyJastrowKernel, JastrowKernelElectronElectronNucleiBase  # noqa: F401
