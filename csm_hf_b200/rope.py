"""RoPE tables for the engine: cos/sin of inv_freq*pos, rounded to bf16.

Follows the un-vendored code the reference runs: llama3 frequency scaling
(transformers modeling_rope_utils.py:550-625, `_compute_llama3_parameters`) and
LlamaRotaryEmbedding.forward (models/llama/modeling_llama.py:122-135): the angle is an
fp32 product, cos/sin are fp32, the cast to the model dtype happens last.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch


def inv_freq(head_dim: int, theta: float, scaling: Optional[dict]) -> torch.Tensor:
    f = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).to(dtype=torch.float) / head_dim))
    kind = (scaling or {}).get("type", (scaling or {}).get("rope_type", "default"))
    if not scaling or kind == "default":
        return f
    if kind != "llama3":
        raise NotImplementedError(f"rope scaling {kind!r} is not on the accelerated path")
    factor, low, high = scaling["factor"], scaling["low_freq_factor"], scaling["high_freq_factor"]
    old = scaling["original_max_position_embeddings"]
    wavelen = 2 * math.pi / f
    scaled = torch.where(wavelen > old / low, f / factor, f)
    smooth = (old / wavelen - low) / (high - low)
    smoothed = (1 - smooth) * scaled / factor + smooth * scaled
    medium = ~(wavelen < old / high) * ~(wavelen > old / low)
    return torch.where(medium, smoothed, scaled)


def tables(head_dim: int, theta: float, scaling: Optional[dict], n_pos: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """[n_pos, head_dim/2] bf16 cos and sin (the second half of HF's table repeats the first)."""
    f = inv_freq(head_dim, theta, scaling)
    ang = torch.arange(n_pos, dtype=torch.float32)[:, None] * f[None, :].to(torch.float32)
    return ang.cos().to(torch.bfloat16).contiguous(), ang.sin().to(torch.bfloat16).contiguous()
