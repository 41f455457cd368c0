from .wf_base import WaveFunction  # noqa: F401
from .slater_jastrow import SlaterJastrow  # noqa: F401
