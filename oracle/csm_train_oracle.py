"""CPU oracle for the TRAINING forward of CSM (loss with labels).  TEST INFRASTRUCTURE ONLY.

Restates CSMModel.forward(labels=...) (reference modeling_csm.py:292-482, loss branch :367-465) with plain torch ops
on top of the step functions of oracle/csm_oracle.py, so that torch autograd of THIS restatement gives reference
gradients for every parameter and for every intermediate a CUDA kernel produces (keep=...).  Pinned against the
reference itself by tests/golden/tiny_train_*.pt (oracle/make_golden.py --train: the imported, unmodified reference's
loss, backbone_loss, decoder_loss and parameter gradients).

Only tests/ and __graft_entry__.smoke() may import this file.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .csm_oracle import CSMOracle, LlamaW, apply_rope, attention, rmsnorm, rope_cos_sin


def llama_train_forward(w: LlamaW, x: torch.Tensor, key_valid: Optional[torch.Tensor], keep: Optional[dict], tag: str):
    """hf:modeling_llama.py:375-425 without a cache: positions 0..S-1, causal (+ padding mask)."""
    B, S, H = x.shape
    cos, sin = rope_cos_sin(w.inv_freq, torch.arange(S), x.dtype)
    h = x
    for l, lw in enumerate(w.layers):
        def k(name, t):
            if keep is not None:
                if t.requires_grad:
                    t.retain_grad()
                keep[f"{tag}.{l}.{name}"] = t
            return t
        r = k("h_in", h)
        hn = k("hn1", rmsnorm(h, lw["ln1"], w.eps))
        q = F.linear(hn, lw["q"]).view(B, S, w.n_heads, w.hd).transpose(1, 2)
        kk = F.linear(hn, lw["k"]).view(B, S, w.n_kv, w.hd).transpose(1, 2)
        v = F.linear(hn, lw["v"]).view(B, S, w.n_kv, w.hd).transpose(1, 2)
        q = apply_rope(q, cos, sin)
        kk = apply_rope(kk, cos, sin)
        # [B,S,(heads+2kv)*hd]: rotated q | rotated k | v, the layout of the engine's qkv rows
        k("qkv", torch.cat([q.transpose(1, 2).reshape(B, S, -1), kk.transpose(1, 2).reshape(B, S, -1),
                            v.transpose(1, 2).reshape(B, S, -1)], dim=-1))
        a = k("attn", attention(q, kk, v, 0, key_valid))
        h = k("h_mid", r + F.linear(a, lw["o"]))
        hn = k("hn2", rmsnorm(h, lw["ln2"], w.eps))
        g = F.linear(hn, lw["gate"])
        u = F.linear(hn, lw["up"])
        act = k("act", F.silu(g) * u)
        h = h + F.linear(act, lw["down"])
    if keep is not None:
        if h.requires_grad:
            h.retain_grad()
        keep[f"{tag}.h_out"] = h
    return rmsnorm(h, w.norm, w.eps)


def training_forward(o: CSMOracle, ids: torch.Tensor, mask: Optional[torch.Tensor], labels: torch.Tensor,
                     keep: Optional[Dict[str, torch.Tensor]] = None):
    """-> (loss, backbone_loss, decoder_loss), modeling_csm.py:367-465."""
    V, NQ = o.V, o.NQ
    B, S = ids.shape[:2]
    h0 = o.embed_sum(ids, mask)                                           # :319-334
    key_valid = None
    if mask is not None:
        fv = mask.sum(dim=-1) > 0                                         # :337-342
        if not bool(fv.all()):
            key_valid = fv
    h = llama_train_forward(o.bb, h0, key_valid, keep, "bb")             # :345-354
    if keep is not None:
        if h.requires_grad:
            h.retain_grad()
        keep["hf"] = h
    c0_all = F.linear(h, o.sd["codebook0_head.weight"])                   # :361
    # codebook-0 cross entropy with the causal shift, computed on float32 logits (:376-389)
    logits = c0_all[:, :-1, :].reshape(-1, V).float()
    shift = labels[:, 1:, 0].reshape(-1)
    backbone_loss = F.cross_entropy(logits, shift, ignore_index=-100)
    # codebooks 1..31 on the frames whose 32 audio labels are all present (:392-399)
    audio_labels = labels[:, :, :NQ]
    idx = (audio_labels != -100).all(dim=2).nonzero(as_tuple=False)
    if idx.numel() == 0:
        decoder_loss = torch.tensor(0.0, dtype=h.dtype)
        return backbone_loss + decoder_loss, backbone_loss, decoder_loss
    fb, ft = idx[:, 0], idx[:, 1]
    frame_h = h[fb, ft - 1]                                               # :405-407 (t = 0 wraps to the last position)
    cbs = ids[fb, ft, :NQ]
    tgt = audio_labels[fb, ft]
    emb = o.sd["audio_embeddings.weight"][(cbs + V * torch.arange(NQ)).view(-1)].view(len(fb), NQ, -1)   # :419-433
    dec_in = torch.cat([F.linear(frame_h, o.sd["projection.weight"]).unsqueeze(1),
                        F.linear(emb, o.sd["projection.weight"])], dim=1)                               # :416,434-443
    if keep is not None:
        if dec_in.requires_grad:
            dec_in.retain_grad()
        keep["dec_in"] = dec_in
    hd = llama_train_forward(o.dec, dec_in, None, keep, "dec")           # :444-447
    if keep is not None:
        if hd.requires_grad:
            hd.retain_grad()
        keep["hdf"] = hd
    cb_hidden = hd[:, 1:NQ, :]                                            # :450-452
    cb_logits = torch.einsum("fcd,cdv->fcv", cb_hidden, o.sd["audio_head"])   # :455-457
    decoder_loss = F.cross_entropy(cb_logits.reshape(-1, V), tgt[:, 1:].reshape(-1), ignore_index=-100)  # :460-467
    return backbone_loss + decoder_loss, backbone_loss, decoder_loss


def loss_and_grads(cfg, state_dict, dtype, ids, mask, labels, keep: Optional[dict] = None):
    """Oracle loss triple and the gradient of `loss` w.r.t. every parameter (autograd of the restatement)."""
    sd = {k: v.to(dtype).clone().requires_grad_(True) for k, v in state_dict.items()}
    o = CSMOracle(cfg, sd, dtype)
    loss, bl, dl = training_forward(o, ids, mask, labels, keep)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)).detach() for k, v in sd.items()}
    return (loss.detach(), bl.detach(), dl.detach()), grads
