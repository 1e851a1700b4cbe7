"""Host-side mirror of the reference's mixer modules (mamba_ssm.modules.*)."""
from .mamba2 import Mamba2  # noqa: F401
from .mamba_simple import Mamba  # noqa: F401
