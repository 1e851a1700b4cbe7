from omnimamba_b200.modules.mamba_simple import Mamba  # noqa: F401
