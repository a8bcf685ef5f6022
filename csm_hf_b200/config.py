"""CSMConfig / CSMOutput: the hyper-parameter and return types of the drop-in boundary.

Mirrors the reference's `CSMConfig` (modeling_csm.py:52-143) and `CSMOutput`
(modeling_csm.py:30-49): same constructor arguments, same attribute names, same
defaults (csm-1b: 16-layer/2048 backbone, 4-layer/1024 decoder, 32 codebooks of 2051
entries, llama3-scaled RoPE).  The sub-configs are plain `LlamaDims` records rather
than `transformers.LlamaConfig` so the product does not need `transformers`; a
`LlamaConfig` object or a dict is accepted wherever the reference accepts one.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field, asdict
from typing import Any, Optional


@dataclass
class LlamaDims:
    """The subset of LlamaConfig the hot path reads (modeling_csm.py:68-109)."""

    hidden_size: int = 2048
    intermediate_size: int = 8192
    num_hidden_layers: int = 16
    num_attention_heads: int = 32
    num_key_value_heads: int = 8
    max_position_embeddings: int = 2048
    rms_norm_eps: float = 1e-5
    rope_theta: float = 500000.0
    rope_scaling: Optional[dict] = field(
        default_factory=lambda: {
            "type": "llama3",
            "factor": 32.0,
            "low_freq_factor": 1.0,
            "high_freq_factor": 4.0,
            "original_max_position_embeddings": 8192,
        }
    )
    vocab_size: int = 128256

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    def to_dict(self) -> dict:
        return asdict(self)

    @staticmethod
    def coerce(obj: Any) -> "LlamaDims":
        """Accept LlamaDims, a dict (config.json) or a transformers.LlamaConfig."""
        if isinstance(obj, LlamaDims):
            return copy.deepcopy(obj)
        if not isinstance(obj, dict):
            src = obj.to_dict() if hasattr(obj, "to_dict") else vars(obj)
        else:
            src = obj
        kw = {}
        for name in (
            "hidden_size intermediate_size num_hidden_layers num_attention_heads "
            "num_key_value_heads max_position_embeddings rms_norm_eps vocab_size"
        ).split():
            if name in src and src[name] is not None:
                kw[name] = src[name]
        # transformers>=5 renames rope_scaling/rope_theta into rope_parameters
        rp = src.get("rope_parameters") or {}
        rs = src.get("rope_scaling") or None
        theta = src.get("rope_theta", None)
        if theta is None:
            theta = rp.get("rope_theta", 500000.0)
        kw["rope_theta"] = float(theta)
        if rs is None and rp and rp.get("rope_type", "default") != "default":
            rs = {k: v for k, v in rp.items() if k != "rope_theta"}
        if rs is not None:
            rs = dict(rs)
            if "type" not in rs and "rope_type" in rs:
                rs["type"] = rs["rope_type"]
        kw["rope_scaling"] = rs
        return LlamaDims(**kw)


def backbone_1b() -> LlamaDims:
    return LlamaDims()


def decoder_100m() -> LlamaDims:
    return LlamaDims(
        hidden_size=1024,
        intermediate_size=8192,
        num_hidden_layers=4,
        num_attention_heads=8,
        num_key_value_heads=2,
        max_position_embeddings=32,
    )


class CSMConfig:
    """Same arguments and attributes as the reference CSMConfig (modeling_csm.py:62-143)."""

    model_type = "csm"

    def __init__(
        self,
        text_vocab_size: int = 128256,
        audio_vocab_size: int = 2051,
        audio_num_codebooks: int = 32,
        max_seq_len: int = 2048,
        backbone_config: Any = None,
        decoder_config: Any = None,
        **kwargs,
    ):
        self.text_vocab_size = text_vocab_size
        self.audio_vocab_size = audio_vocab_size
        self.audio_num_codebooks = audio_num_codebooks
        self.max_seq_len = max_seq_len
        self.backbone_config = LlamaDims.coerce(backbone_config) if backbone_config is not None else backbone_1b()
        self.decoder_config = LlamaDims.coerce(decoder_config) if decoder_config is not None else decoder_100m()
        # the reference overrides these two on both sub-configs (modeling_csm.py:128-129,140-141)
        self.backbone_config.vocab_size = text_vocab_size
        self.backbone_config.max_position_embeddings = max_seq_len
        self.decoder_config.vocab_size = text_vocab_size
        self.decoder_config.max_position_embeddings = audio_num_codebooks
        self.use_return_dict = kwargs.pop("use_return_dict", True)
        self.extra = kwargs

    def to_dict(self) -> dict:
        return {
            "model_type": self.model_type,
            "text_vocab_size": self.text_vocab_size,
            "audio_vocab_size": self.audio_vocab_size,
            "audio_num_codebooks": self.audio_num_codebooks,
            "max_seq_len": self.max_seq_len,
            "backbone_config": self.backbone_config.to_dict(),
            "decoder_config": self.decoder_config.to_dict(),
        }

    @classmethod
    def from_dict(cls, d: dict) -> "CSMConfig":
        d = dict(d)
        d.pop("model_type", None)
        known = {k: d.pop(k) for k in list(d) if k in (
            "text_vocab_size", "audio_vocab_size", "audio_num_codebooks", "max_seq_len",
            "backbone_config", "decoder_config")}
        return cls(**known)

    @classmethod
    def from_reference(cls, ref_cfg: Any) -> "CSMConfig":
        """Build from a reference `CSMConfig` instance (modeling_csm.py:52)."""
        return cls(
            text_vocab_size=ref_cfg.text_vocab_size,
            audio_vocab_size=ref_cfg.audio_vocab_size,
            audio_num_codebooks=ref_cfg.audio_num_codebooks,
            max_seq_len=ref_cfg.max_seq_len,
            backbone_config=ref_cfg.backbone_config,
            decoder_config=ref_cfg.decoder_config,
        )


def tiny_config(**over) -> CSMConfig:
    """A small shape family used by parity tests (same structure, every dimension
    still a multiple of what the kernels tile by)."""
    bb = LlamaDims(hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                   num_attention_heads=4, num_key_value_heads=2)
    dec = LlamaDims(hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                    num_attention_heads=2, num_key_value_heads=1)
    kw = dict(text_vocab_size=512, audio_vocab_size=67, audio_num_codebooks=32, max_seq_len=2048,
              backbone_config=bb, decoder_config=dec)
    kw.update(over)
    return CSMConfig(**kw)


@dataclass
class CSMOutput:
    """Return record of forward / generate_frame (modeling_csm.py:30-49)."""

    last_hidden_state: Any = None
    logits: Any = None
    past_key_values: Any = None
    samples: Any = None
    loss: Any = None
    backbone_loss: Any = None
    decoder_loss: Any = None

    def __getitem__(self, k):
        if isinstance(k, str):
            return getattr(self, k)
        return tuple(v for v in asdict_shallow(self).values() if v is not None)[k]


def asdict_shallow(o) -> dict:
    return {f: getattr(o, f) for f in o.__dataclass_fields__}
