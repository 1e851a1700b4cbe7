"""GatedMLP is imported by the reference (mixer_seq_simple.py:19) but unreachable with d_intermediate=0
(config_mamba.py:7)."""
import torch.nn as nn


class GatedMLP(nn.Module):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("GatedMLP is not on the OmniMamba path (d_intermediate is 0)")
