"""Python mirror of the reference's operator interface (same names, arguments and error behaviour)."""
from .causal_conv1d import causal_conv1d_fn, causal_conv1d_update  # noqa: F401
from .layer_norm import RMSNorm, layer_norm_fn, rms_norm_fn  # noqa: F401
from .layernorm_gated import RMSNorm as RMSNormGated  # noqa: F401
from .layernorm_gated import layernorm_fn, rmsnorm_fn  # noqa: F401
from .selective_scan import selective_scan_fn  # noqa: F401
from .selective_state_update import selective_state_update  # noqa: F401
from .ssd_combined import mamba_chunk_scan_combined, mamba_split_conv1d_scan_combined  # noqa: F401
