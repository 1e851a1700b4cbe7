"""causal_conv1d==1.4.0 surface on libomnissm.so."""
__version__ = "1.4.0+omnissm"
from omnimamba_b200.interface.causal_conv1d import causal_conv1d_fn, causal_conv1d_update  # noqa: F401
