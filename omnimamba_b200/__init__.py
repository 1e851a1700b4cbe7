"""omnimamba_b200 - B200 (sm_100a) kernels for OmniMamba's Mamba-2 selective-scan hot path.

The package exposes the reference's operator surface (``mamba_ssm`` / ``causal_conv1d`` names and
signatures, see ``omnimamba_b200.dropin``) on top of ``libomnissm.so`` (include/omnissm.h), which is
called through ctypes.  There is no CPU or PyTorch fallback: every op raises if the library is
missing or a tensor is not on a CUDA device.
"""
__version__ = "0.1.0"

from . import _cabi  # noqa: F401
from .dropin import install as install_dropin  # noqa: F401
