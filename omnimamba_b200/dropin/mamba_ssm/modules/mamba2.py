from omnimamba_b200.modules.mamba2 import Mamba2  # noqa: F401
