"""Deterministic synthetic weights and contexts for random-init configs.

There are no pretrained weights offline, so every config in BASELINE.json is
"csm-1b random-init".  The reference's own init is not reproducible without the
reference class (HF `post_init` order; `audio_head` is never initialised at all,
modeling_csm.py:236-240), so random-init is defined here instead: every tensor of
the reference `state_dict` (187 keys) gets its own seeded CPU
generator.  The same state_dict is loaded into the reference model when golden
vectors are minted (oracle/make_golden.py), into the oracle, and into the CUDA
engine, so all three see bit-identical parameters.
"""
from __future__ import annotations

import zlib
from typing import Dict, Tuple

import torch

from .config import CSMConfig, LlamaDims


def state_dict_shapes(cfg: CSMConfig) -> Dict[str, Tuple[int, ...]]:
    """Key -> shape of the reference CSMModel.state_dict() (modeling_csm.py:214-245)."""
    shapes: Dict[str, Tuple[int, ...]] = {}

    def llama(prefix: str, d: LlamaDims):
        H, I = d.hidden_size, d.intermediate_size
        hd = d.head_dim
        for l in range(d.num_hidden_layers):
            p = f"{prefix}.layers.{l}"
            shapes[f"{p}.self_attn.q_proj.weight"] = (d.num_attention_heads * hd, H)
            shapes[f"{p}.self_attn.k_proj.weight"] = (d.num_key_value_heads * hd, H)
            shapes[f"{p}.self_attn.v_proj.weight"] = (d.num_key_value_heads * hd, H)
            shapes[f"{p}.self_attn.o_proj.weight"] = (H, d.num_attention_heads * hd)
            shapes[f"{p}.mlp.gate_proj.weight"] = (I, H)
            shapes[f"{p}.mlp.up_proj.weight"] = (I, H)
            shapes[f"{p}.mlp.down_proj.weight"] = (H, I)
            shapes[f"{p}.input_layernorm.weight"] = (H,)
            shapes[f"{p}.post_attention_layernorm.weight"] = (H,)
        shapes[f"{prefix}.norm.weight"] = (H,)

    llama("backbone", cfg.backbone_config)
    llama("decoder", cfg.decoder_config)
    Hb, Hd = cfg.backbone_config.hidden_size, cfg.decoder_config.hidden_size
    shapes["text_embeddings.weight"] = (cfg.text_vocab_size, Hb)
    shapes["audio_embeddings.weight"] = (cfg.audio_vocab_size * cfg.audio_num_codebooks, Hb)
    shapes["projection.weight"] = (Hd, Hb)
    shapes["codebook0_head.weight"] = (cfg.audio_vocab_size, Hb)
    shapes["audio_head"] = (cfg.audio_num_codebooks - 1, Hd, cfg.audio_vocab_size)
    return shapes


def make_state_dict(cfg: CSMConfig, seed: int = 0, dtype: torch.dtype = torch.float32,
                    std: float = 0.02, norm_jitter: float = 0.0,
                    head_gain: float = 1.0, head_pair_gain: float = 0.0) -> Dict[str, torch.Tensor]:
    """N(0, std) for every matrix, norm weights 1 (+ jitter), one generator per key.

    head_gain multiplies codebook0_head / audio_head: the "peaky" variant of
    SURVEY.md §8d, whose top-1 margins are far above one bf16 ulp so that free-running
    greedy tokens are well defined across accumulation orders.

    head_pair_gain > 0 builds "decisive" heads: in every codebook's head two tokens get the
    rows +g*r and -g*r (r = the first token's random row), so one of them wins the argmax by
    a margin of ~2x its own logit unless r.h is within 1/g of zero -- greedy decoding then has
    margins of hundreds of bf16 ulps and free-running ids are comparable exactly across
    implementations (oracle/make_golden.py --decisive; a uniform gain would scale the
    margins and the ulps alike).
    """
    sd: Dict[str, torch.Tensor] = {}
    for name, shape in state_dict_shapes(cfg).items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
        if name.endswith("norm.weight") or name.endswith("layernorm.weight"):
            t = torch.ones(shape, dtype=torch.float32)
            if norm_jitter:
                t += norm_jitter * torch.randn(shape, generator=g, dtype=torch.float32)
        else:
            t = torch.empty(shape, dtype=torch.float32).normal_(0.0, std, generator=g)
            if name in ("codebook0_head.weight", "audio_head"):
                t *= head_gain
                if head_pair_gain > 0:
                    V = cfg.audio_vocab_size
                    if name == "audio_head":                       # [31, H, V]: columns are tokens
                        for c in range(t.shape[0]):
                            b = torch.randperm(V, generator=g)[:2]
                            t[c, :, b[0]] *= head_pair_gain
                            t[c, :, b[1]] = -t[c, :, b[0]]
                    else:                                          # [V, H]: rows are tokens
                        b = torch.randperm(V, generator=g)[:2]
                        t[b[0]] *= head_pair_gain
                        t[b[1]] = -t[b[0]]
        sd[name] = t.to(dtype)
    return sd


def make_context(cfg: CSMConfig, batch: int, frames: int, seed: int = 1234,
                 text_frames: int = 0):
    """Synthetic all-audio context: ids [B,T,33] int64 with the text
    column zero, int32 mask with columns 0..31 set.  `text_frames` > 0 prepends that
    many text-only frames (column 32 = token, mask column 32 only), the shape
    CSMProcessor emits for a prompt (processor.py:254-267)."""
    g = torch.Generator().manual_seed(seed)
    nq = cfg.audio_num_codebooks
    ids = torch.randint(0, cfg.audio_vocab_size, (batch, frames, nq + 1), generator=g, dtype=torch.int64)
    ids[:, :, nq] = 0
    mask = torch.zeros(batch, frames, nq + 1, dtype=torch.int32)
    mask[:, :, :nq] = 1
    if text_frames:
        tt = torch.randint(0, cfg.text_vocab_size, (batch, text_frames), generator=g, dtype=torch.int64)
        ids[:, :text_frames, :] = 0
        ids[:, :text_frames, nq] = tt
        mask[:, :text_frames, :] = 0
        mask[:, :text_frames, nq] = 1
    return ids, mask


def make_padded_context(cfg: CSMConfig, lengths, frames: int, seed: int = 1234, text_frames: int = 0,
                        pad_text_id: int = 0, mask_dtype: torch.dtype = torch.int32):
    """A variable-length batch left-padded to `frames` the way CSMProcessor pads it (processor.py:137-169):
    sequence b keeps its last lengths[b] frames, the frames before them are padding -- ids 0 (text column
    `pad_text_id`), all 33 mask entries 0.  `mask_dtype=torch.float32` reproduces the dtype the processor's
    padding path emits (processor.py:148)."""
    B = len(lengths)
    ids, mask = make_context(cfg, B, frames, seed=seed, text_frames=0)
    nq = cfg.audio_num_codebooks
    g = torch.Generator().manual_seed(seed + 977)
    for b, n in enumerate(lengths):
        if not 1 <= n <= frames:
            raise ValueError("every sequence needs between 1 and `frames` real frames")
        npad = frames - n
        if text_frames:
            tf = min(text_frames, n)
            tt = torch.randint(0, cfg.text_vocab_size, (tf,), generator=g, dtype=torch.int64)
            ids[b, npad:npad + tf, :] = 0
            ids[b, npad:npad + tf, nq] = tt
            mask[b, npad:npad + tf, :] = 0
            mask[b, npad:npad + tf, nq] = 1
        ids[b, :npad, :] = 0
        ids[b, :npad, nq] = pad_text_id
        mask[b, :npad, :] = 0
    return ids, mask.to(mask_dtype)


def make_training_batch(cfg: CSMConfig, batch: int, frames: int, seed: int = 4321, text_frames: int = 2,
                        amortization_ratio: int = 16, pad: int = 0):
    """Synthetic training batch in the shape CSMProcessor emits (processor.py:200-380): `text_frames` text-only frames,
    then audio frames; labels = the audio tokens for codebook 0 on every audio frame and for codebooks 1..31 on one
    frame in `amortization_ratio` (decoder amortisation, processor.py:340-360), -100 elsewhere.  `pad` > 0 left-pads
    sequence 0 with that many all-zero-mask frames whose labels are -100 (processor.py:142-160).
    -> ids [B,S,33] int64, mask [B,S,33] int32, labels [B,S,33] int64."""
    g = torch.Generator().manual_seed(seed)
    nq = cfg.audio_num_codebooks
    ids, mask = make_context(cfg, batch, frames, seed=seed, text_frames=text_frames)
    labels = torch.full_like(ids, -100)
    labels[:, text_frames:, 0] = ids[:, text_frames:, 0]
    n_audio = frames - text_frames
    for b in range(batch):
        n_sel = max(1, n_audio // amortization_ratio)
        sel = torch.randperm(n_audio, generator=g)[:n_sel] + text_frames
        labels[b, sel, :nq] = ids[b, sel, :nq]
    if pad:
        ids[0, :pad] = 0
        mask[0, :pad] = 0
        labels[0, :pad] = -100
    return ids, mask, labels
