"""Short generate() used as the ncu target: csm-1b, --ctx context frames, --frames new frames, --batch sequences."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from csm_hf_b200.config import CSMConfig  # noqa: E402
from csm_hf_b200.modeling import CSMModel  # noqa: E402
from csm_hf_b200.synthetic import make_context, make_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--ctx", type=int, default=2048)
ap.add_argument("--frames", type=int, default=6)
ap.add_argument("--reps", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda", 0)
cfg = CSMConfig()
model = CSMModel(cfg, make_state_dict(cfg, seed=0, dtype=torch.bfloat16), device=dev, max_batch=a.batch,
                 max_ctx=a.ctx + a.frames + 8)
ids, mask = make_context(cfg, a.batch, a.ctx)
ids, mask = ids.to(dev), mask.to(dev)
for _ in range(a.reps):
    out = model.generate(ids, mask, max_new_frames=a.frames, temperature=0, stop_on_all_zeros=False)
torch.cuda.synchronize()
ms, n = model.last_decode_ms()
print(f"B={a.batch} ctx={a.ctx}: decode {ms / max(n, 1):.3f} ms/frame over {n} frames; launches {model.engine().info(4)}")
