from .metropolis import Metropolis, SamplerBase  # noqa: F401
from .ensemble import Walkers  # noqa: F401
