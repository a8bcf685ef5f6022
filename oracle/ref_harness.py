"""Drive the UNMODIFIED reference (/root/reference/modeling_csm.py) on CPU.

TEST / MEASUREMENT INFRASTRUCTURE: used by oracle/make_golden.py to mint tests/golden/*.pt in
the build container, and by bench.py's reference arm (`--impl reference`, cpu_baseline), which
times the reference's own CPU path from the copy oracle/stage_ref.py stages under oracle/_ref/
(git-ignored; /root/reference itself does not exist on the GPU box).  No test under -m gpu and
nothing in the product imports this module.

Caveats handled here, each established by probes in SURVEY.md §0 / §8c:
  1. transformers>=5 turns the reference's all-ones [B,1] decode mask into "attend to
     position 0 only"; the shim below drops all-ones 2-D masks, restoring the 4.49
     semantics the reference was written against (fact 4);
  2. `audio_head` is torch.empty in the reference (fact 3) -> always loaded from our
     state_dict;
  3. greedy == topk=1, temperature=1.0 (fact 1);
  4. ties: `_multinomial_sample_one_no_sync` is replaced by argmax (lowest index) so
     that greedy is a function (fact 2).
"""
from __future__ import annotations

import os
import sys
import warnings

import torch

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")   # oracle/stage_ref.py (git-ignored copy)
REF_DIR = os.environ.get("CSM_REFERENCE_DIR", "/root/reference")
if not os.path.isfile(os.path.join(REF_DIR, "modeling_csm.py")) and os.path.isfile(os.path.join(_STAGED, "modeling_csm.py")):
    REF_DIR = _STAGED          # the GPU box: /root/reference does not exist there, the staged copy travels with the repo


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "modeling_csm.py"))


def load_reference_module():
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import modeling_csm as M  # noqa: the reference, imported unchanged
    return M


def build_reference_model(cfg, state_dict, dtype=torch.float32, canonical_ties=True):
    """cfg: csm_hf_b200.config.CSMConfig; returns the reference CSMModel in eval mode."""
    M = load_reference_module()
    from transformers import LlamaConfig

    def to_llama(d):
        return LlamaConfig(
            vocab_size=d.vocab_size, hidden_size=d.hidden_size, intermediate_size=d.intermediate_size,
            num_hidden_layers=d.num_hidden_layers, num_attention_heads=d.num_attention_heads,
            num_key_value_heads=d.num_key_value_heads, max_position_embeddings=d.max_position_embeddings,
            rms_norm_eps=d.rms_norm_eps, attention_dropout=0.0, rope_theta=d.rope_theta,
            rope_scaling=dict(d.rope_scaling) if d.rope_scaling else None, hidden_act="silu",
        )

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rcfg = M.CSMConfig(
            text_vocab_size=cfg.text_vocab_size, audio_vocab_size=cfg.audio_vocab_size,
            audio_num_codebooks=cfg.audio_num_codebooks, max_seq_len=cfg.max_seq_len,
            backbone_config=to_llama(cfg.backbone_config), decoder_config=to_llama(cfg.decoder_config),
        )
        model = M.CSMModel(rcfg)
    missing, unexpected = model.load_state_dict(state_dict, strict=False)
    missing = [k for k in missing if "rotary_emb" not in k]
    assert not missing and not unexpected, (missing, unexpected)
    model = model.to(dtype).eval()

    orig = model.backbone.forward

    def shim(*a, attention_mask=None, **kw):  # caveat 1
        if attention_mask is not None and attention_mask.dim() == 2 and bool((attention_mask == 1).all()):
            attention_mask = None
        return orig(*a, attention_mask=attention_mask, **kw)

    model.backbone.forward = shim
    if canonical_ties:  # caveat 4 (name is looked up at call time, modeling_csm.py:189)
        M._multinomial_sample_one_no_sync = lambda p: p.argmax(-1, keepdim=True).to(torch.int)
    return model


@torch.inference_mode()
def reference_generate(model, ids, mask, n_frames):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return model.generate(ids, mask, max_new_frames=n_frames, temperature=1.0, topk=1,
                              use_cache=True, stop_on_all_zeros=False)


@torch.inference_mode()
def reference_trace(model, ids, mask, n_frames):
    """Free-running greedy generate through the reference's own generate_frame, also
    recording what the reference computes at each sampling point: last_h, c0 logits and
    the 31 per-codebook logits (captured by wrapping sample_topk)."""
    M = load_reference_module()
    captured = []
    orig_sample = M.sample_topk

    def spy(logits, topk, temperature):
        captured.append(logits.detach().clone())
        return orig_sample(logits, topk, temperature)

    M.sample_topk = spy
    try:
        frames, traces = [], []
        kv = None
        run_ids, run_mask = ids, mask
        B = ids.shape[0]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for _ in range(n_frames):
                captured.clear()
                out = model.generate_frame(run_ids, run_mask, temperature=1.0, topk=1,
                                           past_key_values=kv, use_cache=True, return_dict=True)
                kv = out.past_key_values
                traces.append({
                    "last_h": out.last_hidden_state.clone(),
                    "c0_logits": out.logits.clone(),
                    "cb_logits": torch.stack(captured[1:], dim=1),
                })
                frames.append(out.samples)
                run_ids = torch.cat([out.samples, torch.zeros(B, 1, dtype=torch.long)], dim=1).unsqueeze(1)
                run_mask = torch.zeros(B, 1, 33, dtype=mask.dtype)
                run_mask[:, :, :32] = 1
        return torch.stack(frames, dim=1), traces
    finally:
        M.sample_topk = orig_sample
