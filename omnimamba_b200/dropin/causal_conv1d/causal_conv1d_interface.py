from omnimamba_b200.interface.causal_conv1d import CausalConv1dFn, causal_conv1d_fn, causal_conv1d_update  # noqa: F401
