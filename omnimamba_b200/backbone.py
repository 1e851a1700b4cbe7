"""Host-side mirror of the caller of the hot path: the 48-layer loop around `Mamba2` that OmniMamba runs
(/root/reference/models/stage2/mixer_seq_simple.py:404-437: fused add+norm -> mixer, per layer, then the final norm;
block.py:86-117), the LoRA-wrapped `in_proj` it installs on every mixer (lora.py:76-106, 263-279: r=8, alpha=32,
dropout 0.05, one adapter per task) and the stage-1 t2i embedding / head / shifted cross-entropy around it
(models/omnimamba.py:252-280, mamba_vlm.py:88-100).

This is what bench.py's whole-model workloads (BASELINE.json configs 2, 3, 4) and the stack parity tests run: random-init
weights of the reference architecture, the reference's parameter names, our kernels underneath.  It is NOT a copy of the
reference model zoo: tokenizer, vision towers, VQ decoder, adaLN conditioning and the mmu projector are out of scope
(SURVEY.md 8)."""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .interface.layer_norm import RMSNorm, layer_norm_fn
from .modules.mamba2 import Mamba2


class LoRALinear(nn.Linear):
    """`in_proj` as the reference's _find_and_replace leaves it (lora.py:76-106): a frozen-or-not base weight plus one
    rank-r adapter per task, result = x W^T + B_task(A_task(dropout(x))) * alpha / r (lora.py:263-279)."""

    def __init__(self, in_features, out_features, r=8, lora_alpha=32, lora_dropout=0.05, bias=False, device=None, dtype=None):
        super().__init__(in_features, out_features, bias=bias, device=device, dtype=dtype)
        self.r, self.lora_alpha, self.scaling = r, lora_alpha, lora_alpha / r
        self.lora_dropout = nn.Dropout(lora_dropout) if lora_dropout > 0 else nn.Identity()
        self.task_types = "t2i"
        kw = dict(bias=False, device=device, dtype=dtype)
        for task in ("mmu", "t2i"):
            a, b = nn.Linear(in_features, r, **kw), nn.Linear(r, out_features, **kw)
            nn.init.kaiming_uniform_(a.weight, a=math.sqrt(5))
            nn.init.zeros_(b.weight)
            setattr(self, f"{task}_lora_A0", a)
            setattr(self, f"{task}_lora_B0", b)

    def forward(self, x):
        from .interface.gemm import lora_linear
        a = getattr(self, f"{self.task_types}_lora_A0").weight
        b = getattr(self, f"{self.task_types}_lora_B0").weight
        return lora_linear(x, self.weight, self.bias, a, b, self.scaling, self.lora_dropout)


class Block(nn.Module):
    """Add -> norm -> mixer with the fused add+norm kernel (block.py:86-117, adaLN / MLP branches unused by stage 1)."""

    def __init__(self, d_model, layer_idx, ssm_cfg=None, norm_epsilon=1e-5, residual_in_fp32=True, lora=True, device=None, dtype=None):
        super().__init__()
        self.residual_in_fp32 = residual_in_fp32
        self.norm = RMSNorm(d_model, eps=norm_epsilon, device=device, dtype=dtype)
        self.mixer = Mamba2(d_model, layer_idx=layer_idx, device=device, dtype=dtype, **(ssm_cfg or {}))
        if lora:
            base = self.mixer.in_proj
            wrapped = LoRALinear(base.in_features, base.out_features, bias=base.bias is not None, device=device, dtype=dtype)
            wrapped.weight = base.weight
            self.mixer.in_proj = wrapped
        self.layer_idx = layer_idx

    def forward(self, hidden_states, residual=None, inference_params=None):
        hidden_states, residual = layer_norm_fn(hidden_states, self.norm.weight, self.norm.bias, residual=residual, prenorm=True,
                                                residual_in_fp32=self.residual_in_fp32, eps=self.norm.eps, is_rms_norm=True)
        return self.mixer(hidden_states, inference_params=inference_params), residual

    def allocate_inference_cache(self, batch_size, max_seqlen, dtype=None, **kw):
        return self.mixer.allocate_inference_cache(batch_size, max_seqlen, dtype=dtype, **kw)


class MixerStack(nn.Module):
    """n_layer Blocks + final norm: the loop of MixerModel.forward (mixer_seq_simple.py:404-437) on given embeddings."""

    def __init__(self, d_model=2048, n_layer=48, ssm_cfg=None, norm_epsilon=1e-5, residual_in_fp32=True, lora=True, device=None,
                 dtype=None):
        super().__init__()
        self.residual_in_fp32 = residual_in_fp32
        self.layers = nn.ModuleList([Block(d_model, i, ssm_cfg, norm_epsilon, residual_in_fp32, lora, device, dtype)
                                     for i in range(n_layer)])
        self.norm_f = RMSNorm(d_model, eps=norm_epsilon, device=device, dtype=dtype)
        for layer in self.layers:  # GPT-2 residual scaling of out_proj (mixer_seq_simple.py:_init_weights)
            nn.init.kaiming_uniform_(layer.mixer.out_proj.weight, a=math.sqrt(5))
            with torch.no_grad():
                layer.mixer.out_proj.weight /= math.sqrt(n_layer)

    def set_lora_mode(self, task="t2i"):
        for layer in self.layers:
            if hasattr(layer.mixer.in_proj, "task_types"):
                layer.mixer.in_proj.task_types = task

    def allocate_inference_cache(self, batch_size, max_seqlen, dtype=None, **kw):
        return {i: l.allocate_inference_cache(batch_size, max_seqlen, dtype=dtype, **kw) for i, l in enumerate(self.layers)}

    def forward(self, hidden_states, inference_params=None):
        residual = None
        for layer in self.layers:
            hidden_states, residual = layer(hidden_states, residual, inference_params=inference_params)
        return layer_norm_fn(hidden_states, self.norm_f.weight, self.norm_f.bias, eps=self.norm_f.eps, residual=residual,
                             prenorm=False, residual_in_fp32=self.residual_in_fp32, is_rms_norm=True)


class T2IModel(nn.Module):
    """Stage-1 text-to-image training graph around the stack (models/omnimamba.py:252-280): image-token and caption
    embeddings, caption MLP, learned positions, 48 layers, img_head, shifted cross-entropy over the 256 image tokens."""

    def __init__(self, d_model=2048, n_layer=48, vocab_size=50288, vqvae_vocab_size=16384, num_tokens=256, caption_len=73,
                 device=None, dtype=None):
        super().__init__()
        kw = dict(device=device, dtype=dtype)
        self.img_embeddings = nn.Embedding(vqvae_vocab_size, d_model, **kw)
        self.embedding = nn.Embedding(vocab_size, d_model, **kw)
        self.caption_embed = nn.Sequential(nn.Linear(d_model, d_model, **kw), nn.GELU(approximate="tanh"),
                                           nn.Linear(d_model, d_model, **kw))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_tokens + caption_len, d_model, **kw))
        nn.init.trunc_normal_(self.pos_embed, 0.0, 0.02)
        nn.init.normal_(self.img_embeddings.weight, std=0.02)
        nn.init.normal_(self.embedding.weight, std=0.02)
        self.backbone = MixerStack(d_model, n_layer, ssm_cfg=dict(), **kw)
        self.img_head = nn.Linear(d_model, vqvae_vocab_size, bias=False, **kw)

    def freeze_backbones(self, stage="align"):
        """models/omnimamba.py:119-146: stage "align" trains only the embeddings, the caption MLP, the positions, the image
        head and the LoRA adapters; "finetune" trains everything."""
        if stage == "finetune":
            self.requires_grad_(True)
            return self
        assert stage == "align"
        self.requires_grad_(False)
        for m in (self.img_embeddings, self.embedding, self.caption_embed, self.img_head):
            m.requires_grad_(True)
        self.pos_embed.requires_grad_(True)
        for name, p in self.backbone.named_parameters():
            if "lora" in name.lower():
                p.requires_grad_(True)
        return self

    def embed(self, image_ids, caption_ids):
        img = self.img_embeddings(image_ids)
        txt = self.caption_embed(self.embedding(caption_ids))
        x = torch.cat((txt[:, :-1], img, txt[:, -1:]), dim=1)
        return x + self.pos_embed[:, :x.shape[1]]

    def labels(self, image_ids, caption_ids, ignore_id=-100):
        b = image_ids.shape[0]
        pad = lambda n: torch.full((b, n), ignore_id, device=image_ids.device, dtype=torch.long)
        return torch.cat([pad(caption_ids.shape[1] - 1), image_ids.long(), pad(1)], dim=1)

    def forward(self, image_ids, caption_ids):
        from .interface.linear_ce import linear_cross_entropy
        self.backbone.set_lora_mode("t2i")
        h = self.backbone(self.embed(image_ids, caption_ids))
        labels = self.labels(image_ids, caption_ids)
        # shifted CE (mamba_vlm.py:96-100): position t predicts label t+1; the head GEMM and the loss are one op
        return linear_cross_entropy(h[:, :-1].reshape(-1, h.shape[-1]), self.img_head.weight, labels[:, 1:].reshape(-1))


class InferenceParams:
    """Duck type of models/stage2/generation.py:19-36."""

    def __init__(self, max_seqlen, max_batch_size, seqlen_offset=0, batch_size_offset=0, key_value_memory_dict=None,
                 lengths_per_sample=None):
        self.max_seqlen, self.max_batch_size = max_seqlen, max_batch_size
        self.seqlen_offset, self.batch_size_offset = seqlen_offset, batch_size_offset
        self.key_value_memory_dict = {} if key_value_memory_dict is None else key_value_memory_dict
        self.lengths_per_sample = lengths_per_sample

    def reset(self, max_seqlen, max_batch_size):
        self.max_seqlen, self.max_batch_size, self.seqlen_offset = max_seqlen, max_batch_size, 0
        if self.lengths_per_sample is not None:
            self.lengths_per_sample.zero_()
