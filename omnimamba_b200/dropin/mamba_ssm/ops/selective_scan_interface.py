from omnimamba_b200.interface.selective_scan import (  # noqa: F401
    SelectiveScanFn, mamba_inner_fn, selective_scan_fn, selective_scan_ref)
