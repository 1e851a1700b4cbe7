"""Checkpoint helpers used only by MambaLMHeadModel.from_pretrained (mixer_seq_simple.py:527-532): local files
only - there is no hub access here."""
import json
import os

import torch


def load_config_hf(model_name):
    with open(os.path.join(model_name, "config.json")) as f:
        return json.load(f)


def load_state_dict_hf(model_name, device=None, dtype=None):
    """`model_name` is a local directory holding `model.safetensors` or `pytorch_model.bin` (tensors only: weights_only)."""
    st = os.path.join(model_name, "model.safetensors")
    if os.path.exists(st):
        from safetensors.torch import load_file   # (ships with transformers)
        sd = load_file(st, device="cpu")
    else:
        sd = torch.load(os.path.join(model_name, "pytorch_model.bin"), map_location="cpu", weights_only=True)
    if dtype is not None:
        sd = {k: v.to(dtype=dtype) for k, v in sd.items()}
    return {k: v.to(device=device) for k, v in sd.items()}
