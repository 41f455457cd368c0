from .metropolis import Metropolis, SamplerBase  # noqa: F401
from .ensemble import Walkers  # noqa: F401
from .gradient_samplers import GeneralizedMetropolis, Hamiltonian  # noqa: F401
