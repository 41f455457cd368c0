"""qmctorch_b200 - B200-native (sm_100a) walker-parallel Slater-Jastrow hot path behind the
QMCTorch API: ``from qmctorch_b200.wavefunction import SlaterJastrow``,
``from qmctorch_b200.sampler import Metropolis``, ``from qmctorch_b200.solver import Solver``."""
from .utils import set_torch_double_precision  # noqa: F401

__version__ = "0.1.0"
