from .jastrow_factor_electron_nuclei import JastrowFactorElectronNuclei  # noqa: F401
from .jastrow_factor_electron_nuclei import JastrowFactorElectronNuclei as JastrowFactor  # noqa: F401
from .kernels import PadeJastrowKernel, JastrowKernelElectronNucleiBase  # noqa: F401
