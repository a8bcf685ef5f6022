"""Mint tests/golden/*.pt by running the UNMODIFIED reference on CPU.

Build-container only (needs /root/reference).  Usage:
    python oracle/make_golden.py [--full]

Every fixture stores the recipe (config name, weight seed/jitter, context seed/shape,
dtype) next to the reference's outputs, so tests rebuild the exact inputs from
csm_hf_b200.synthetic and compare against what the reference produced:
  frames      int64 [B,n,32]   reference greedy tokens (topk=1, canonical ties)
  last_h      [n,B,H]          CSMOutput.last_hidden_state per frame
  c0_logits   [n,B,V]          CSMOutput.logits per frame
  cb_logits   [n,B,31,V]       logits the reference passed to sample_topk for codebooks 1..31
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from csm_hf_b200.config import CSMConfig, tiny_config  # noqa: E402
from csm_hf_b200.synthetic import make_context, make_state_dict  # noqa: E402
from oracle import ref_harness as R  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def mint(name, cfg_name, cfg, dtype, B, T, n, wseed, jitter, cseed, text_frames, keep_cb=True):
    sd = make_state_dict(cfg, seed=wseed, norm_jitter=jitter)
    ids, mask = make_context(cfg, B, T, seed=cseed, text_frames=text_frames)
    t0 = time.time()
    model = R.build_reference_model(cfg, sd, dtype)
    frames, tr = R.reference_trace(model, ids, mask, n)
    frames2 = R.reference_generate(model, ids, mask, n)
    assert torch.equal(frames, frames2), "reference generate() != generate_frame() loop"
    out = {
        "recipe": dict(config=cfg_name, dtype=str(dtype).split(".")[-1], batch=B, ctx_frames=T, new_frames=n,
                       weight_seed=wseed, norm_jitter=jitter, ctx_seed=cseed, text_frames=text_frames),
        "frames": frames,
        "last_h": torch.stack([t["last_h"] for t in tr]),
        "c0_logits": torch.stack([t["c0_logits"] for t in tr]),
    }
    if keep_cb:
        out["cb_logits"] = torch.stack([t["cb_logits"] for t in tr])
    os.makedirs(GOLD, exist_ok=True)
    torch.save(out, os.path.join(GOLD, name))
    print(f"{name}: frames {tuple(frames.shape)} in {time.time() - t0:.1f}s; c0 of frame0 = {frames[:, 0, 0].tolist()}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="also mint the csm-1b config #1 fixtures (6 GB, minutes)")
    a = ap.parse_args()
    assert R.reference_available(), "needs /root/reference"
    torch.manual_seed(0)
    tiny = tiny_config()
    for dt, tag in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
        mint(f"tiny_{tag}.pt", "tiny", tiny, dt, B=2, T=6, n=4, wseed=0, jitter=0.1, cseed=1234, text_frames=2)
        mint(f"tiny_b1_{tag}.pt", "tiny", tiny, dt, B=1, T=16, n=8, wseed=3, jitter=0.1, cseed=77, text_frames=0)
    if a.full:
        full = CSMConfig()
        # BASELINE.json configs[0]: csm-1b random-init, greedy, 16-frame context, 8 frames, batch 1
        mint("csm1b_cfg1_fp32.pt", "csm-1b", full, torch.float32, B=1, T=16, n=8, wseed=0, jitter=0.0,
             cseed=1234, text_frames=0, keep_cb=False)
        mint("csm1b_cfg1_bf16.pt", "csm-1b", full, torch.bfloat16, B=1, T=16, n=8, wseed=0, jitter=0.0,
             cseed=1234, text_frames=0, keep_cb=True)


if __name__ == "__main__":
    main()
