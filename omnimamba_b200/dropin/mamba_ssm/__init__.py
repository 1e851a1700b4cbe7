"""mamba_ssm==2.2.2 surface on libomnissm.so (only what hustvl/OmniMamba reaches)."""
__version__ = "2.2.2+omnissm"
from mamba_ssm.ops.selective_scan_interface import selective_scan_fn, mamba_inner_fn  # noqa: F401
from mamba_ssm.modules.mamba_simple import Mamba  # noqa: F401
from mamba_ssm.modules.mamba2 import Mamba2  # noqa: F401
