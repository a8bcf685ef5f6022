"""CSMModel: host-side mirror of the reference class (modeling_csm.py:192-702) over the CUDA engine.

Same method names, argument meaning and return types as the reference for the generation
path (`generate`, `generate_frame`, inference `forward`, `setup_caches`, `reset_caches`,
`_embed_audio`, `_embed_tokens`), same state_dict keys.  PyTorch is used for device memory
and streams only; every arithmetic step of the path runs in libcsm_b200.so (ctypes,
include/csm_b200.h).  There is no CPU or eager fallback: without the library or a B200 the
constructor of the engine raises.

`CSMModel` is an `nn.Module` whose parameter tree has the reference's attribute names
(`backbone.layers[i].self_attn.q_proj.weight`, `text_embeddings.weight`, `audio_head`, ... --
modeling_csm.py:222-240) and therefore the reference's 187 `state_dict()` keys; the sub-modules hold
parameters only (their arithmetic lives in the engine), so `.to()`, `.parameters()`,
`load_state_dict()` and `state_dict()` behave as on the reference.

Differences from the reference, all deliberate and documented in DESIGN.md:
  * `temperature == 0` (or the reference's own spelling, `topk == 1`) is greedy decoding with
    ties broken to the lowest index (the reference breaks them randomly);
  * stochastic top-k sampling draws with counter-based noise seeded from `torch.initial_seed()`
    instead of torch's global generator: same distribution, different stream;
  * a float32 attention mask (what CSMProcessor emits when it pads, processor.py:148) is accepted with a
    bf16 model (the reference raises a dtype error there);
  * `past_key_values` is an opaque handle to the engine's in-place KV cache;
  * `labels` (training): csm_hf_b200/training.py -- loss and parameter gradients from csm_train_step.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Dict, Optional

import torch

from . import native, rope
from .config import CSMConfig, CSMOutput
from .synthetic import state_dict_shapes

_LAYER_KEYS = ["self_attn.q_proj.weight", "self_attn.k_proj.weight", "self_attn.v_proj.weight",
               "self_attn.o_proj.weight", "mlp.gate_proj.weight", "mlp.up_proj.weight", "mlp.down_proj.weight",
               "input_layernorm.weight", "post_attention_layernorm.weight"]


class KVHandle:
    """What `past_key_values` is on this path: a token for the engine-owned cache.  It is valid for the next
    call only: the cache it names is overwritten in place (the reference returns a new DynamicCache state
    each step, modeling_csm.py:653,659)."""

    def __init__(self, model: "CSMModel", length: int, batch: int, serial: int):
        self.model, self.length, self.batch, self.serial = model, length, batch, serial

    def get_seq_length(self) -> int:
        return self.length


class _Params(torch.nn.Module):
    """A node of the reference's module tree that only holds parameters (`backbone`, `layers[i]`, `self_attn`,
    `q_proj`, ... -- modeling_csm.py:156-167,222-240).  Indexable where the reference has a ModuleList."""

    def __getitem__(self, i):
        return getattr(self, str(i))

    def __len__(self):
        return len(self._modules)

    def __iter__(self):
        return iter(self._modules.values())

    def forward(self, *a, **k):
        raise NotImplementedError("sub-modules of the B200 CSMModel hold parameters only; call generate / "
                                  "generate_frame / forward on the model (the arithmetic runs in libcsm_b200.so)")


class _Engine:
    """Owns one CsmCtx (packed weights + workspace) on one device."""

    def __init__(self, cfg: CSMConfig, sd: Dict[str, torch.Tensor], device: torch.device, max_batch: int, max_ctx: int):
        if device.type != "cuda":
            raise RuntimeError("CSMModel generation runs only on a CUDA (B200, sm_100a) device")
        self.lib = native.load()
        self.device, self.max_batch, self.max_ctx = device, max_batch, max_ctx
        self.cfg = cfg
        keep = []  # keep host tables alive during csm_create

        def llama_shape(d, n_pos):
            cos, sin = rope.tables(d.head_dim, d.rope_theta, d.rope_scaling, n_pos)
            keep.extend([cos, sin])
            s = native.LlamaShape()
            s.hidden, s.inter, s.layers = d.hidden_size, d.intermediate_size, d.num_hidden_layers
            s.heads, s.kv_heads, s.eps = d.num_attention_heads, d.num_key_value_heads, d.rms_norm_eps
            s.rope_cos, s.rope_sin, s.n_pos = cos.data_ptr(), sin.data_ptr(), n_pos
            return s

        sh = native.Shapes()
        sh.text_vocab, sh.audio_vocab, sh.n_codebooks = cfg.text_vocab_size, cfg.audio_vocab_size, cfg.audio_num_codebooks
        sh.backbone = llama_shape(cfg.backbone_config, max_ctx)
        sh.decoder = llama_shape(cfg.decoder_config, 32)

        def dev(name):
            t = sd[name]
            if t.device != device or t.dtype != torch.bfloat16 or not t.is_contiguous():
                t = t.to(device=device, dtype=torch.bfloat16).contiguous()
            keep.append(t)
            return t.data_ptr()

        def layer_ptrs(prefix, n):
            arr = (C.c_void_p * (n * native.W_PER_LAYER))()
            for l in range(n):
                for j, k in enumerate(_LAYER_KEYS):
                    arr[l * native.W_PER_LAYER + j] = dev(f"{prefix}.layers.{l}.{k}")
            return arr

        w = native.Weights()
        w.text_embeddings = dev("text_embeddings.weight")
        w.audio_embeddings = dev("audio_embeddings.weight")
        w.projection = dev("projection.weight")
        w.codebook0_head = dev("codebook0_head.weight")
        w.audio_head = dev("audio_head")
        w.backbone_norm = dev("backbone.norm.weight")
        w.decoder_norm = dev("decoder.norm.weight")
        bl = layer_ptrs("backbone", cfg.backbone_config.num_hidden_layers)
        dl = layer_ptrs("decoder", cfg.decoder_config.num_hidden_layers)
        w.backbone_layers = C.cast(bl, C.POINTER(C.c_void_p))
        w.decoder_layers = C.cast(dl, C.POINTER(C.c_void_p))
        self.ctx = C.c_void_p()
        with torch.cuda.device(device):
            torch.cuda.synchronize()
            rc = self.lib.csm_create(C.byref(sh), C.byref(w), max_batch, max_ctx, self._stream(), C.byref(self.ctx))
            try:
                native.check(self.lib, self.ctx, rc)
            except Exception:
                self.close()
                raise
            torch.cuda.synchronize()

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "ctx", None) is not None and self.ctx:
            self.lib.csm_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def call(self, fn, *args):
        return native.check(self.lib, self.ctx, fn(self.ctx, *args))

    def info(self, what: int) -> int:
        return int(self.lib.csm_info(self.ctx, what))


class CSMModel(torch.nn.Module):
    """Drop-in for the reference CSMModel on the generation path."""

    config_class = CSMConfig
    base_model_prefix = "csm"

    def __init__(self, config: CSMConfig, state_dict: Optional[Dict[str, torch.Tensor]] = None,
                 device: Optional[torch.device] = None, max_batch: int = 1, max_ctx: Optional[int] = None):
        super().__init__()
        self.config = config
        self._device = torch.device(device) if device is not None else torch.device("cuda", 0)
        self._engine: Optional[_Engine] = None
        self._max_batch = max_batch
        self._max_ctx = max_ctx or max(4096, config.max_seq_len + 512)
        self._using_kv_cache = False
        self._kv: Optional[KVHandle] = None
        self._kv_serial = 0
        self._loaded = False
        if state_dict is not None:
            self.load_state_dict(state_dict)

    # ------------------------------------------------------------------ weights
    @property
    def device(self) -> torch.device:
        return self._device

    @property
    def dtype(self) -> torch.dtype:
        return torch.bfloat16

    def _set_param(self, key: str, t: torch.Tensor):
        node = self
        parts = key.split(".")
        for name in parts[:-1]:
            if name not in node._modules:
                node.add_module(name, _Params())
            node = node._modules[name]
        node._parameters[parts[-1]] = torch.nn.Parameter(t, requires_grad=False)

    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], strict: bool = True, assign: bool = False):
        """The reference's 187 keys; tensors are converted to bf16 on the model's device."""
        want = state_dict_shapes(self.config)
        missing = [k for k in want if k not in state_dict]
        unexpected = [k for k in state_dict if k not in want and "rotary_emb" not in k]
        if strict and (missing or unexpected):
            raise RuntimeError(f"state_dict mismatch: missing {missing[:4]}, unexpected {unexpected[:4]}")
        for k, shape in want.items():
            if k in state_dict:
                t = state_dict[k]
                if tuple(t.shape) != tuple(shape):
                    raise RuntimeError(f"size mismatch for {k}: {tuple(t.shape)} vs {tuple(shape)}")
                self._set_param(k, t.detach().to(device=self._device, dtype=torch.bfloat16).contiguous())
        self._loaded = len(dict(self.named_parameters())) == len(want)
        self._drop_engine()
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def _apply(self, fn, *a, **k):
        """.to() / .cuda() / .bfloat16(): move the parameters, then rebuild the engine lazily on the new device."""
        out = super()._apply(fn, *a, **k)
        for p in self.parameters():
            self._device = p.device
            break
        self._drop_engine()
        return out

    @classmethod
    def from_reference(cls, ref_model, device=None, **kw) -> "CSMModel":
        """Wrap an instantiated reference `CSMModel` (same weights, bf16 on `device`)."""
        cfg = CSMConfig.from_reference(ref_model.config)
        sd = {k: v for k, v in ref_model.state_dict().items() if "rotary_emb" not in k}
        return cls(cfg, sd, device=device, **kw)

    @classmethod
    def from_pretrained(cls, path: str, torch_dtype=torch.bfloat16, device=None, **kw) -> "CSMModel":
        """Local directory with config.json + model.safetensors / pytorch_model.bin in the
        reference's 187-key layout (train.py:370-379).  No hub download (no network)."""
        if torch_dtype not in (None, torch.bfloat16):
            raise ValueError("the B200 engine computes in bf16; pass torch_dtype=torch.bfloat16")
        with open(os.path.join(path, "config.json")) as f:
            cfg = CSMConfig.from_dict(json.load(f))
        st = os.path.join(path, "model.safetensors")
        if os.path.isfile(st):
            from safetensors.torch import load_file
            sd = load_file(st)
        else:
            sd = torch.load(os.path.join(path, "pytorch_model.bin"), map_location="cpu", weights_only=True)
        return cls(cfg, sd, device=device, **kw)

    def save_pretrained(self, path: str):
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, "config.json"), "w") as f:
            json.dump(self.config.to_dict(), f, indent=1)
        from safetensors.torch import save_file
        save_file({k: v.detach().cpu().contiguous() for k, v in self.state_dict().items()},
                  os.path.join(path, "model.safetensors"))

    # ------------------------------------------------------------------ engine
    def _drop_engine(self):
        if getattr(self, "_engine", None) is not None:
            self._engine.close()
        self._engine, self._kv = None, None
        self._kv_serial = getattr(self, "_kv_serial", 0) + 1     # handles issued by the old engine are dead

    def engine(self, batch: int = 1, ctx_len: int = 0) -> _Engine:
        need_b, need_t = max(batch, self._max_batch), max(ctx_len, self._max_ctx)
        if need_b > 32:
            raise ValueError(f"batch {need_b} > 32 sequences per GPU: shard the batch (csm_hf_b200.dist.generate_sharded)")
        e = self._engine
        if e is None or e.max_batch < need_b or e.max_ctx < need_t:
            if not self._loaded:
                raise RuntimeError("CSMModel has no weights: load_state_dict / from_pretrained first")
            self._drop_engine()
            self._max_batch, self._max_ctx = need_b, need_t
            self._engine = _Engine(self.config, dict(self.state_dict()), self.device, need_b, need_t)
        return self._engine

    def setup_caches(self, max_batch_size: int):
        """modeling_csm.py:284-286 -- here it also sizes the engine's KV cache."""
        self._using_kv_cache = True
        self._max_batch = max(self._max_batch, int(max_batch_size))

    def reset_caches(self):
        """modeling_csm.py:288-290."""
        if self._engine is not None:
            self._engine.call(self._engine.lib.csm_reset)
        self._kv = None
        self._kv_serial += 1

    # ------------------------------------------------------------------ small gathers (API parity; not hot)
    def _embed_audio(self, codebook: int, tokens: torch.Tensor) -> torch.Tensor:
        """modeling_csm.py:247-259."""
        return self.audio_embeddings.weight[tokens.to(self.device) + codebook * self.config.audio_vocab_size]

    def _embed_tokens(self, tokens: torch.Tensor) -> torch.Tensor:
        """modeling_csm.py:261-282 -> [B,S,33,H]."""
        tokens = tokens.to(self.device)
        nq, V = self.config.audio_num_codebooks, self.config.audio_vocab_size
        text = self.text_embeddings.weight[tokens[:, :, -1]].unsqueeze(-2)
        aud = self.audio_embeddings.weight[tokens[:, :, :-1] + V * torch.arange(nq, device=self.device)]
        return torch.cat([aud, text], dim=-2)

    def embed_sum(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor]) -> torch.Tensor:
        """Fused K1 kernel: sum over the 33 slots of mask*embedding (modeling_csm.py:327-334) -> [B,S,H] bf16."""
        ids, mask, _ = self._prep_inputs(input_ids, attention_mask, check=False, none_is_ones=False)
        B, S = ids.shape[:2]
        e = self.engine(B, 0)
        out = torch.empty(B, S, self.config.backbone_config.hidden_size, dtype=torch.bfloat16, device=self.device)
        e.call(e.lib.csm_embed_sum, ids.data_ptr(), mask.data_ptr() if mask is not None else None, B, S,
               out.data_ptr(), e._stream())
        return out

    # ------------------------------------------------------------------ input checks
    def _prep_inputs(self, input_ids, attention_mask, check=True, none_is_ones=True):
        """-> (ids int64, mask int32, padded?) on the device.  attention_mask=None means "all 33 slots present", as
        in the reference (modeling_csm.py:328-332: the mask multiply is skipped); the engine's own NULL convention
        (audio slots only) is used for the decode rows generate() builds itself, never for a caller's None."""
        nq = self.config.audio_num_codebooks
        if input_ids.dim() != 3 or input_ids.shape[-1] != nq + 1:
            raise ValueError(f"input_ids must be [B, S, {nq + 1}]")
        if input_ids.dtype not in (torch.int64, torch.int32):
            raise ValueError("input_ids must be an integer tensor")
        mask = attention_mask
        if mask is not None:
            if mask.shape != input_ids.shape:
                raise ValueError("attention_mask must have the shape of input_ids")
            if mask.dtype.is_floating_point:
                # CSMProcessor pads with a float32 zero mask (processor.py:148); the mask is a 0/1 indicator
                if check and not bool(((mask == 0) | (mask == 1)).all()):
                    raise ValueError("a floating-point attention_mask must hold 0 / 1 only")
                mask = mask != 0
        elif none_is_ones:
            mask = torch.ones(input_ids.shape, dtype=torch.int32, device=input_ids.device)
        padded = False
        if check:
            V, TV = self.config.audio_vocab_size, self.config.text_vocab_size
            m = torch.ones_like(input_ids) if mask is None else (mask != 0)
            a, t = input_ids[..., :nq], input_ids[..., nq]
            bad = ((a < 0) | (a >= V)) & (m[..., :nq] != 0)
            badt = ((t < 0) | (t >= TV)) & (m[..., nq] != 0)
            if bool(bad.any()) or bool(badt.any()):
                raise IndexError("token id out of range of the embedding tables")
            if mask is not None:
                padded = not bool((mask != 0).any(dim=-1).all())
        ids = input_ids.to(device=self.device, dtype=torch.int64).contiguous()
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.int32).contiguous()
        return ids, mask, padded

    def _set_sampling(self, e, temperature, topk, seq_base: int = 0):
        """sample_topk's arguments (modeling_csm.py:179-189) -> engine sampling mode.  Greedy when
        temperature == 0 or topk == 1; otherwise top-k + temperature sampling, reproducible for a given
        torch.manual_seed() and call sequence."""
        if temperature is None or topk is None or temperature < 0 or topk < 1:
            raise ValueError(f"invalid sampling arguments temperature={temperature}, topk={topk}")
        if temperature == 0 or topk == 1:
            e.call(e.lib.csm_set_sampling, 1, 1.0, 0, 0)
            return
        self._sample_calls = getattr(self, "_sample_calls", 0) + 1
        seed = (int(torch.initial_seed()) * 0x9E3779B97F4A7C15 + self._sample_calls) & 0xFFFFFFFFFFFFFFFF
        e.call(e.lib.csm_set_sampling, int(topk), float(temperature), seed, int(seq_base))

    # ------------------------------------------------------------------ generate_frame / forward
    def generate_frame(self, input_ids, attention_mask, position_ids=None, temperature=1.0, topk=50,
                       past_key_values=None, use_cache=None, output_attentions=None, output_hidden_states=None,
                       return_dict=None, force_tokens: Optional[torch.Tensor] = None, return_codebook_logits=False):
        """modeling_csm.py:484-589.  `force_tokens` [B,32] (extension) teacher-forces the decoder;
        `return_codebook_logits` adds `.codebook_logits` [B,31,V] to the output."""
        if position_ids is not None:
            raise NotImplementedError("explicit position_ids are not supported (the reference passes None)")
        if output_attentions or output_hidden_states:
            raise NotImplementedError("output_attentions / output_hidden_states are not available on the fused path")
        return_dict = True if return_dict is None else return_dict
        ids, mask, padded = self._prep_inputs(input_ids, attention_mask)
        B, S = ids.shape[:2]
        e = self.engine(B, 0)
        self._set_sampling(e, temperature, topk, getattr(self, "seq_base", 0))
        if past_key_values is None:
            e.call(e.lib.csm_reset)          # a call without a cache starts a new context
        elif not isinstance(past_key_values, KVHandle) or past_key_values.model is not self:
            raise ValueError("past_key_values must be the handle returned by this model's generate_frame")
        else:
            kv = past_key_values
            if kv.serial != self._kv_serial or kv.batch != B or kv.length != e.call(e.lib.csm_cache_len):
                raise ValueError("stale past_key_values: the engine's cache has moved on since this handle was issued "
                                 "(another generate / generate_frame / reset_caches call, or a different batch size)")
            if padded:
                raise NotImplementedError("padding is only honoured in the first (prefill) call of a context, as in "
                                          "the reference's generate()")
        start = e.call(e.lib.csm_cache_len)
        if start + S > e.max_ctx:
            e = self._grow_ctx(e, start + S)
        H, V = self.config.backbone_config.hidden_size, self.config.audio_vocab_size
        samples = torch.empty(B, 32, dtype=torch.int64, device=self.device)
        last_h = torch.empty(B, H, dtype=torch.bfloat16, device=self.device)
        c0 = torch.empty(B, V, dtype=torch.bfloat16, device=self.device)
        cb = torch.empty(B, 31, V, dtype=torch.bfloat16, device=self.device) if return_codebook_logits else None
        ft = None
        if force_tokens is not None:
            ft = force_tokens.to(device=self.device, dtype=torch.int64).contiguous()
        e.call(e.lib.csm_generate_frame, ids.data_ptr(), mask.data_ptr() if mask is not None else None, B, S,
               ft.data_ptr() if ft is not None else None, samples.data_ptr(), last_h.data_ptr(), c0.data_ptr(),
               cb.data_ptr() if cb is not None else None, e._stream())
        self._kv_serial += 1
        self._kv = KVHandle(self, start + S, B, self._kv_serial)
        if input_ids.device.type != "cuda":
            samples = samples.to(input_ids.device)
        if not return_dict:
            return samples
        out = CSMOutput(last_hidden_state=last_h, logits=c0, past_key_values=self._kv, samples=samples)
        if cb is not None:
            out.codebook_logits = cb
        return out

    def _grow_ctx(self, e, need):
        raise ValueError(f"context of {need} positions exceeds max_ctx={e.max_ctx}; construct CSMModel with a larger max_ctx")

    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None, use_cache=None,
                output_attentions=None, output_hidden_states=None, return_dict=None, temperature=1.0, topk=50,
                generate_frame=False, labels=None):
        """Inference branch of modeling_csm.py:292-365,467-482: last_hidden_state + codebook-0 logits."""
        if labels is not None:   # training: loss branch of modeling_csm.py:367-465, forward + backward on the GPU
            from .training import training_forward
            return training_forward(self, input_ids, attention_mask, labels, return_dict)
        out = self.generate_frame(input_ids, attention_mask, position_ids=position_ids, temperature=0.0, topk=1,
                                  past_key_values=past_key_values, use_cache=use_cache, return_dict=True)
        return_dict = True if return_dict is None else return_dict
        if not return_dict:
            return (out.last_hidden_state, out.logits, out.past_key_values)
        return CSMOutput(last_hidden_state=out.last_hidden_state, logits=out.logits, past_key_values=out.past_key_values)


    # ------------------------------------------------------------------ generate
    def generate(self, input_ids: torch.Tensor, attention_mask: torch.Tensor, max_new_frames: int = 100,
                 temperature: float = 1.0, topk: int = 50, use_cache: bool = True, stop_on_all_zeros: bool = True,
                 reserve_frames: int = 0):
        """modeling_csm.py:591-702 -> LongTensor [B, n, 32] on the device of `input_ids`.

        CPU inputs take the host-buffer C-ABI call (csm_generate_host: pinned H2D copy, all
        frames on the device, one D2H copy); CUDA inputs are consumed in place."""
        if not use_cache:
            raise NotImplementedError("use_cache=False loses the context in the reference itself (SURVEY.md fact 6)")
        ids, mask, _ = self._prep_inputs(input_ids, attention_mask) if input_ids.device.type == "cuda" else (None, None, None)
        B, T = input_ids.shape[:2]
        e = self.engine(B, T + max(max_new_frames, reserve_frames))   # (reserve_frames: room for generate_more)
        self._set_sampling(e, temperature, topk, getattr(self, "seq_base", 0))
        if max_new_frames <= 0:
            return torch.zeros(B, 0, 32, dtype=torch.long, device=input_ids.device)
        if input_ids.device.type == "cuda":
            frames = torch.empty(B, max_new_frames, 32, dtype=torch.int64, device=self.device)
            e.call(e.lib.csm_generate, ids.data_ptr(), mask.data_ptr() if mask is not None else None, B, T,
                   max_new_frames, int(bool(stop_on_all_zeros)), frames.data_ptr(), e._stream())
            n = e.call(e.lib.csm_frames_done, e._stream())
        else:
            self._prep_inputs(input_ids, attention_mask)   # validation only (host tensors)
            hi = input_ids.to(torch.int64).contiguous()
            am = attention_mask if attention_mask is not None else torch.ones(input_ids.shape, dtype=torch.int32)
            hm = (am != 0).to(torch.int32).contiguous() if am.dtype.is_floating_point else am.to(torch.int32).contiguous()
            if not hi.is_pinned():
                hi = hi.pin_memory()
            if hm is not None and not hm.is_pinned():
                hm = hm.pin_memory()
            frames = torch.empty(B, max_new_frames, 32, dtype=torch.int64).pin_memory()
            n_out = C.c_int(0)
            e.call(e.lib.csm_generate_host, hi.data_ptr(), hm.data_ptr() if hm is not None else None, B, T,
                   max_new_frames, int(bool(stop_on_all_zeros)), frames.data_ptr(), C.byref(n_out), e._stream())
            n = n_out.value
        self._kv = None
        self._kv_serial += 1
        return frames[:, :n].contiguous()

    def generate_more(self, batch: int, n_more: int, stop_on_all_zeros: bool = True) -> torch.Tensor:
        """Continue the last generate() for up to `n_more` frames (same decode steps, issued as a further chunk):
        -> LongTensor [B, n, 32] on the model's device.  Used by batch-sharded generation to exchange the stop flag
        between chunks (csm_hf_b200.dist.generate_sharded)."""
        e = self._engine
        if e is None:
            raise RuntimeError("generate_more needs a preceding generate()")
        if n_more <= 0:
            return torch.zeros(batch, 0, 32, dtype=torch.long, device=self.device)
        frames = torch.empty(batch, n_more, 32, dtype=torch.int64, device=self.device)
        e.call(e.lib.csm_generate_more, batch, n_more, int(bool(stop_on_all_zeros)), frames.data_ptr(), e._stream())
        n = e.call(e.lib.csm_frames_done, e._stream())
        self._kv_serial += 1
        return frames[:, :n].contiguous()

    def last_decode_ms(self):
        """(ms, frames): device time of the decode-frame launches of the last generate()."""
        e = self._engine
        ms, n = C.c_float(0), C.c_int(0)
        e.call(e.lib.csm_last_decode_ms, C.byref(ms), C.byref(n))
        return ms.value, n.value

    def set_stepped(self, on: bool):
        e = self.engine()
        e.call(e.lib.csm_set_stepped, int(on))
