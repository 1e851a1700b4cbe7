"""MHA is imported by the reference (mixer_seq_simple.py:18) but unreachable with attn_layer_idx=[]
(config_mamba.py:17): constructing it is out of scope."""
import torch.nn as nn


class MHA(nn.Module):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("attention layers are not on the OmniMamba path (attn_layer_idx is empty)")
