from .solver import Solver  # noqa: F401
from . import distributed  # noqa: F401
