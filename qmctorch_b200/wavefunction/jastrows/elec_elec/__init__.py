from .jastrow_factor_electron_electron import JastrowFactorElectronElectron  # noqa: F401
from .jastrow_factor_electron_electron import JastrowFactorElectronElectron as JastrowFactor  # noqa: F401
from .kernels import PadeJastrowKernel, JastrowKernelElectronElectronBase  # noqa: F401
