from omnimamba_b200.interface.layer_norm import LayerNormFn, RMSNorm, layer_norm_fn, rms_norm_fn  # noqa: F401
