"""Shared parity helpers (tests only)."""
import os

import torch

from csm_hf_b200.config import CSMConfig, tiny_config
from csm_hf_b200.synthetic import make_context, make_padded_context, make_state_dict

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    g = torch.load(os.path.join(GOLD, name), weights_only=False)
    r = g["recipe"]
    cfg = tiny_config() if r["config"] == "tiny" else CSMConfig()
    dtype = getattr(torch, r["dtype"])
    sd = make_state_dict(cfg, seed=r["weight_seed"], norm_jitter=r["norm_jitter"], head_pair_gain=r.get("head_pair_gain", 0.0))
    if r.get("lengths"):
        ids, mask = make_padded_context(cfg, r["lengths"], r["ctx_frames"], seed=r["ctx_seed"], text_frames=r["text_frames"])
    else:
        ids, mask = make_context(cfg, r["batch"], r["ctx_frames"], seed=r["ctx_seed"], text_frames=r["text_frames"])
    return g, cfg, dtype, sd, ids, mask


def top2_margin(logits):
    """top-1 minus top-2 value along the last dim (float32)."""
    v = torch.topk(logits.float(), 2, dim=-1).values
    return v[..., 0] - v[..., 1]


def assert_logits_close(got, want, rel, what):
    """max |got-want| <= rel * max|want| -- the tolerance for bf16 pipelines that share
    rounding points but not accumulation order."""
    got, want = got.float(), want.float()
    scale = want.abs().max().item()
    err = (got - want).abs().max().item()
    assert err <= rel * scale, f"{what}: max err {err:.5f} > {rel} * {scale:.4f}"
    return err / max(scale, 1e-30)


def assert_tokens_match_where_decided(got_tok, want_tok, want_logits, tol_abs, what):
    """Greedy ids must be identical wherever the reference's own top-1/top-2 margin
    exceeds 2*tol_abs (elsewhere the argmax is numerically undecided)."""
    decided = top2_margin(want_logits) > 2 * tol_abs
    bad = (got_tok != want_tok) & decided
    assert not bool(bad.any()), f"{what}: {int(bad.sum())} decided tokens differ"
    return float(decided.float().mean())


def dense_grad(entry):
    """A gradient of tests/golden/tiny_train_bf16.pt: embedding tables are stored as (rows, values)."""
    if isinstance(entry, dict):
        g = torch.zeros(entry["shape"], dtype=entry["values"].dtype)
        g[entry["rows"]] = entry["values"]
        return g
    return entry
