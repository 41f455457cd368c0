from .metropolis import Metropolis, SamplerBase  # noqa: F401
from .walkers import Walkers  # noqa: F401
