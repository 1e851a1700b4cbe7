from omnimamba_b200.interface.ssd_combined import (  # noqa: F401
    MambaChunkScanCombinedFn, MambaSplitConv1dScanCombinedFn, mamba_chunk_scan_combined,
    mamba_split_conv1d_scan_combined)
