"""Drop-in `mamba_ssm` / `causal_conv1d` packages backed by libomnissm.so.

The reference imports these names unmodified (/root/reference/models/stage2/mixer_seq_simple.py:15-20,30 and
models/stage2/block.py:10).  `install()` puts this directory at the FRONT of sys.path so `import mamba_ssm`
and `import causal_conv1d` resolve here; equivalently add `<repo>/omnimamba_b200/dropin` to PYTHONPATH."""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def install() -> str:
    for name in ("mamba_ssm", "causal_conv1d"):
        mod = sys.modules.get(name)
        if mod is not None and not os.path.abspath(getattr(mod, "__file__", "") or "").startswith(_HERE):
            raise RuntimeError(f"{name} is already imported from {getattr(mod, '__file__', '?')}; "
                               "call omnimamba_b200.install_dropin() before importing the model code")
    if _HERE not in sys.path:
        sys.path.insert(0, _HERE)
    return _HERE
