"""CPU oracle for CSM two-stage frame generation.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it.  The CUDA engine
in csm_hf_b200/ never routes through it.

It is a plain-torch (CPU) restatement of the reference algorithm, one explicit
function per arithmetic step, each citing the reference line it follows:

  /root/reference/modeling_csm.py  (the CSM wrapper: embed-sum, heads, sampling,
                                    generate_frame, generate)
  hf: = transformers/models/llama/modeling_llama.py, modeling_rope_utils.py,
        cache_utils.py — the un-vendored third-party code (pinned 4.49.0 in the
        reference's requirements.txt:4, 5.5.0 installed) where the Llama arithmetic
        lives.

Parity pinning: the reference ships no tests or golden vectors, so
this oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, run in the build
container by oracle/make_golden.py (which imports /root/reference/modeling_csm.py
unchanged) and committed under tests/golden/.  tests/test_oracle_golden.py checks
this file against those fixtures on every CPU run.

Semantics fixed here:
  * greedy == reference `topk=1, temperature=1.0`, ties broken towards the lowest
    index (the reference breaks ties randomly, modeling_csm.py:170-189);
  * decode steps attend to every cached position (transformers-4.49 behaviour of the
    all-ones [B,1] mask the reference passes, modeling_csm.py:337-342,684-690);
  * rounding points in bf16 mode follow the reference op by op.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- RoPE
def llama3_inv_freq(head_dim: int, theta: float, scaling: Optional[dict]) -> torch.Tensor:
    """hf:modeling_rope_utils.py:550-625 (_compute_llama3_parameters), fp32."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).to(dtype=torch.float) / head_dim))
    if not scaling or scaling.get("type", scaling.get("rope_type", "default")) == "default":
        return inv_freq
    factor = scaling["factor"]
    low = scaling["low_freq_factor"]
    high = scaling["high_freq_factor"]
    old = scaling["original_max_position_embeddings"]
    low_wl = old / low
    high_wl = old / high
    wavelen = 2 * math.pi / inv_freq
    inv_l = torch.where(wavelen > low_wl, inv_freq / factor, inv_freq)
    smooth = (old / wavelen - low) / (high - low)
    smoothed = (1 - smooth) * inv_l / factor + smooth * inv_l
    medium = ~(wavelen < high_wl) * ~(wavelen > low_wl)
    return torch.where(medium, smoothed, inv_l)


def rope_cos_sin(inv_freq: torch.Tensor, positions: torch.Tensor, dtype: torch.dtype):
    """hf:modeling_llama.py:122-135: freqs = inv_freq*pos in fp32, cat, cos/sin, cast.
    positions: int64 [S] -> cos, sin [S, head_dim] in `dtype`."""
    freqs = positions.to(torch.float32)[:, None] * inv_freq[None, :].to(torch.float32)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    """hf:modeling_llama.py:138-142."""
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """hf:modeling_llama.py:146-168; x [B, heads, S, hd], cos/sin [S, hd].
    In bf16 each of x*cos, rot(x)*sin and the sum rounds to bf16."""
    return (x * cos[None, None]) + (rotate_half(x) * sin[None, None])


# --------------------------------------------------------------------------- blocks
def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """hf:modeling_llama.py:62-67: fp32 normalise, cast to input dtype, then * weight."""
    dt = x.dtype
    xf = x.to(torch.float32)
    var = xf.pow(2).mean(-1, keepdim=True)
    xf = xf * torch.rsqrt(var + eps)
    return w * xf.to(dt)


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, q_start: int,
              key_valid: Optional[torch.Tensor] = None) -> torch.Tensor:
    """hf:integrations/sdpa_attention.py:40-104 restated: GQA by head repetition, scale
    hd^-1/2, fp32 softmax; query i (absolute position q_start+i) sees keys <= that
    position -- causal in prefill, everything cached in a decode step.
    key_valid [B,T] bool (padded batches, modeling_csm.py:337-342 -> hf:masking_utils.py
    padding mask): keys of all-zero-mask frames are hidden; a query that sees no key at
    all gets a zero output (torch >= 2.5 SDPA "safe softmax", which is what the reference
    runs on in this container).
    q [B,Hq,S,hd]; k,v [B,Hkv,T,hd] -> [B,S,Hq*hd]."""
    B, Hq, S, hd = q.shape
    Hkv, T = k.shape[1], k.shape[2]
    rep = Hq // Hkv
    kf = k.to(torch.float32).repeat_interleave(rep, dim=1)
    vf = v.to(torch.float32).repeat_interleave(rep, dim=1)
    s = torch.matmul(q.to(torch.float32), kf.transpose(-1, -2)) * (hd ** -0.5)
    qpos = torch.arange(q_start, q_start + S)[:, None]
    kpos = torch.arange(T)[None, :]
    hidden = (kpos > qpos)[None, None].expand(B, 1, S, T)
    if key_valid is not None:
        hidden = hidden | ~key_valid[:, None, None, :T].bool()
    s = s.masked_fill(hidden, float("-inf"))
    p = torch.softmax(s, dim=-1)
    p = torch.where(hidden.all(dim=-1, keepdim=True), torch.zeros_like(p), p)   # fully hidden rows -> 0, not NaN
    o = torch.matmul(p, vf).to(q.dtype)
    return o.transpose(1, 2).reshape(B, S, Hq * hd)


@dataclass
class LlamaW:
    """One Llama stack's weights, pulled from the reference state_dict by prefix."""

    H: int
    I: int
    L: int
    n_heads: int
    n_kv: int
    eps: float
    inv_freq: torch.Tensor
    layers: List[Dict[str, torch.Tensor]] = field(default_factory=list)
    norm: torch.Tensor = None

    @property
    def hd(self) -> int:
        return self.H // self.n_heads


def _llama_from_sd(sd, prefix: str, dims) -> LlamaW:
    w = LlamaW(H=dims.hidden_size, I=dims.intermediate_size, L=dims.num_hidden_layers,
               n_heads=dims.num_attention_heads, n_kv=dims.num_key_value_heads,
               eps=dims.rms_norm_eps,
               inv_freq=llama3_inv_freq(dims.hidden_size // dims.num_attention_heads,
                                        dims.rope_theta, dims.rope_scaling))
    for l in range(w.L):
        p = f"{prefix}.layers.{l}."
        w.layers.append({
            "q": sd[p + "self_attn.q_proj.weight"], "k": sd[p + "self_attn.k_proj.weight"],
            "v": sd[p + "self_attn.v_proj.weight"], "o": sd[p + "self_attn.o_proj.weight"],
            "gate": sd[p + "mlp.gate_proj.weight"], "up": sd[p + "mlp.up_proj.weight"],
            "down": sd[p + "mlp.down_proj.weight"],
            "ln1": sd[p + "input_layernorm.weight"], "ln2": sd[p + "post_attention_layernorm.weight"],
        })
    w.norm = sd[prefix + ".norm.weight"]
    return w


class KVCache:
    """Pre-sized stand-in for DynamicCache (hf:cache_utils.py:88-133): append == write at `len`."""

    def __init__(self, L, B, n_kv, cap, hd, dtype):
        self.k = torch.zeros(L, B, n_kv, cap, hd, dtype=dtype)
        self.v = torch.zeros(L, B, n_kv, cap, hd, dtype=dtype)
        self.len = 0


def llama_forward(w: LlamaW, x: torch.Tensor, cache: KVCache, trace: Optional[dict] = None,
                  key_valid: Optional[torch.Tensor] = None) -> torch.Tensor:
    """hf:modeling_llama.py:375-425 (LlamaModel.forward) with inputs_embeds=x [B,S,H];
    positions continue from the cache length (:394-397); appends S positions to `cache`.
    key_valid [B, start+S]: padding mask of a padded prefill (None: nothing hidden)."""
    B, S, H = x.shape
    start = cache.len
    pos = torch.arange(start, start + S)
    cos, sin = rope_cos_sin(w.inv_freq, pos, x.dtype)
    h = x
    for l, lw in enumerate(w.layers):
        # hf:modeling_llama.py:303-332 (LlamaDecoderLayer.forward)
        r = h
        hn = rmsnorm(h, lw["ln1"], w.eps)
        q = F.linear(hn, lw["q"]).view(B, S, w.n_heads, w.hd).transpose(1, 2)
        k = F.linear(hn, lw["k"]).view(B, S, w.n_kv, w.hd).transpose(1, 2)
        v = F.linear(hn, lw["v"]).view(B, S, w.n_kv, w.hd).transpose(1, 2)
        q = apply_rope(q, cos, sin)
        k = apply_rope(k, cos, sin)
        cache.k[l, :, :, start:start + S] = k
        cache.v[l, :, :, start:start + S] = v
        a = attention(q, cache.k[l, :, :, :start + S], cache.v[l, :, :, :start + S], start, key_valid)
        h = r + F.linear(a, lw["o"])
        r = h
        hn = rmsnorm(h, lw["ln2"], w.eps)
        g = F.linear(hn, lw["gate"])
        u = F.linear(hn, lw["up"])
        h = r + F.linear(F.silu(g) * u, lw["down"])  # hf:modeling_llama.py:182-184
        if trace is not None:
            trace.setdefault("layer_out", []).append(h.clone())
    cache.len = start + S
    return rmsnorm(h, w.norm, w.eps)  # hf:modeling_llama.py:421


def greedy(logits: torch.Tensor) -> torch.Tensor:
    """modeling_csm.py:179-189 at topk=1, temperature=1: the kept set is the maxima;
    canonical tie-break = lowest index (torch.argmax returns the first maximum)."""
    return torch.argmax(logits, dim=-1)


# --------------------------------------------------------------------------- model
class CSMOracle:
    def __init__(self, cfg, state_dict: Dict[str, torch.Tensor], dtype: torch.dtype = torch.float32):
        self.cfg = cfg
        self.dtype = dtype
        sd = {k: v.to(dtype) for k, v in state_dict.items()}
        self.sd = sd
        self.bb = _llama_from_sd(sd, "backbone", cfg.backbone_config)
        self.dec = _llama_from_sd(sd, "decoder", cfg.decoder_config)
        self.V = cfg.audio_vocab_size
        self.NQ = cfg.audio_num_codebooks

    # modeling_csm.py:247-282 + :328-334: 33-way offset gather, mask multiply, sum over slots
    def embed_sum(self, ids: torch.Tensor, mask: Optional[torch.Tensor]) -> torch.Tensor:
        text = self.sd["text_embeddings.weight"][ids[:, :, -1]].unsqueeze(-2)
        aud_ids = ids[:, :, :-1] + self.V * torch.arange(self.NQ)
        aud = self.sd["audio_embeddings.weight"][aud_ids]
        e = torch.cat([aud, text], dim=-2)                       # [B,S,33,H]
        if mask is not None:
            e = e * mask.unsqueeze(-1)          # int mask keeps the dtype
        return e.sum(dim=2)

    def embed_audio(self, codebook: int, tok: torch.Tensor) -> torch.Tensor:
        """modeling_csm.py:247-259."""
        return self.sd["audio_embeddings.weight"][tok + codebook * self.V]

    def new_cache(self, B: int, cap: int) -> KVCache:
        b = self.bb
        return KVCache(b.L, B, b.n_kv, cap, b.hd, self.dtype)

    # modeling_csm.py:292-365 inference branch
    def backbone_step(self, ids, mask, cache: KVCache, trace=None):
        h = self.embed_sum(ids, mask)
        if trace is not None:
            trace["embed"] = h.clone()
        # modeling_csm.py:337-342: frames whose 33 mask entries are all zero are padding.  The 2-D mask reaches
        # the backbone only in the prefill call: a decode step's [B,1] all-ones mask hides nothing, so cached
        # padding positions (K = V = 0) ARE attended to from then on (SURVEY.md fact 8 -- the reference's quirk).
        key_valid = None
        if mask is not None and ids.shape[1] > 1 and cache.len == 0:
            fv = mask.sum(dim=-1) > 0
            if not bool(fv.all()):
                key_valid = fv
        hs = llama_forward(self.bb, h, cache, trace, key_valid)
        last_h = hs[:, -1, :]
        c0_logits = F.linear(last_h, self.sd["codebook0_head.weight"])   # :361 (last position only)
        return last_h, c0_logits

    # modeling_csm.py:484-589
    def generate_frame(self, ids, mask, cache: KVCache, trace: Optional[dict] = None,
                       force_tokens: Optional[torch.Tensor] = None):
        """force_tokens [B,32]: teacher forcing -- the decoder is fed these tokens instead
        of its own argmax, so per-codebook logits of two implementations stay comparable
        even when a near-tie flips a sample."""
        B = ids.shape[0]
        last_h, c0_logits = self.backbone_step(ids, mask, cache, trace)
        toks = torch.zeros(B, self.NQ, dtype=torch.long)
        c0 = greedy(c0_logits)
        toks[:, 0] = c0
        fed = c0 if force_tokens is None else force_tokens[:, 0]
        d = self.dec
        dcache = KVCache(d.L, B, d.n_kv, self.NQ + 1, d.hd, self.dtype)
        curr = torch.stack([last_h, self.embed_audio(0, fed)], dim=1)      # :535-536
        hdec = llama_forward(d, F.linear(curr, self.sd["projection.weight"]), dcache)  # :542-552
        cb_logits = []
        for i in range(1, self.NQ):                                         # :555-576
            li = torch.matmul(hdec[:, -1, :], self.sd["audio_head"][i - 1])
            cb_logits.append(li)
            ci = greedy(li)
            toks[:, i] = ci
            if i < self.NQ - 1:
                fed = ci if force_tokens is None else force_tokens[:, i]
                e = self.embed_audio(i, fed).unsqueeze(1)
                hdec = llama_forward(d, F.linear(e, self.sd["projection.weight"]), dcache)
        if trace is not None:
            trace["last_h"] = last_h.clone()
            trace["c0_logits"] = c0_logits.clone()
            trace["cb_logits"] = torch.stack(cb_logits, dim=1)             # [B,31,V]
        return toks, last_h, c0_logits

    # modeling_csm.py:591-702
    def generate(self, ids, mask, max_new_frames: int, stop_on_all_zeros: bool = False,
                 traces: Optional[list] = None, force_frames: Optional[torch.Tensor] = None):
        B, T = ids.shape[:2]
        cache = self.new_cache(B, T + max_new_frames)
        frames = []
        run_ids, run_mask = ids, mask
        for f in range(max_new_frames):
            tr = {} if traces is not None else None
            force = None if force_frames is None else force_frames[:, f]
            toks, _, _ = self.generate_frame(run_ids, run_mask, cache, tr, force)
            if traces is not None:
                traces.append(tr)
            if stop_on_all_zeros and bool(torch.all(toks == 0)):          # :662-663
                break
            frames.append(toks)
            nxt = toks if force is None else force
            run_ids = torch.cat([nxt, torch.zeros(B, 1, dtype=torch.long)], dim=1).unsqueeze(1)  # :675-680
            run_mask = torch.zeros(B, 1, self.NQ + 1, dtype=mask.dtype)
            run_mask[:, :, : self.NQ] = 1                                   # :684-687
        if frames:
            return torch.stack(frames, dim=1)
        return torch.zeros(B, 0, self.NQ, dtype=torch.long)
