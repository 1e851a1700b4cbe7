from omnimamba_b200.interface.layernorm_gated import (  # noqa: F401
    LayerNorm, LayerNormFn, RMSNorm, layernorm_fn, rmsnorm_fn)
