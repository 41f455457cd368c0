from .orbital_configurations import OrbitalConfigurations  # noqa: F401
from .slater_pooling import SlaterPooling  # noqa: F401
