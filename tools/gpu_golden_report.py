"""Report (no asserts) the engine's relative errors against the reference-minted fixtures of the benchmarked
configurations, per frame and quantity: max |err| / max |ref|, mean |err| / max |ref|, decided-id mismatches."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import load_golden, top2_margin  # noqa: E402
from csm_hf_b200.modeling import CSMModel  # noqa: E402

dev = torch.device("cuda", 0)
names = sys.argv[1:] or ["csm1b_t2048_b1_bf16.pt", "csm1b_t256_b8_bf16.pt", "csm1b_t256_b32_bf16.pt"]
for name in names:
    g, cfg, dtype, sd, ids, mask = load_golden(name)
    B, T = ids.shape[:2]
    model = CSMModel(cfg, sd, device=dev, max_batch=B, max_ctx=T + 16)
    rows = g["recipe"].get("cb_rows")
    kv, run_ids, run_mask = None, ids, mask
    for f in range(g["frames"].shape[1]):
        out = model.generate_frame(run_ids, run_mask, temperature=0, past_key_values=kv, force_tokens=g["frames"][:, f],
                                   return_codebook_logits=True)
        kv = out.past_key_values
        cb = out.codebook_logits.cpu()
        cb = cb if rows is None else cb[rows]
        for what, got, want in (("last_h", out.last_hidden_state.cpu(), g["last_h"][f]), ("c0", out.logits.cpu(), g["c0_logits"][f]),
                                ("cb", cb, g["cb_logits"][f])):
            d = (got.float() - want.float()).abs()
            sc = float(want.float().abs().max())
            print(f"{name} f{f} {what:7s} max {float(d.max()) / sc:.4f} mean {float(d.mean()) / sc:.5f} range {sc:.3f} nan {bool(torch.isnan(got.float()).any())}")
        tol = 0.06 * float(g["c0_logits"][f].float().abs().max())
        dec = top2_margin(g["c0_logits"][f]) > 2 * tol
        bad = (out.samples.cpu()[:, 0] != g["frames"][:, f, 0]) & dec
        print(f"   c0 ids: decided {int(dec.sum())}/{dec.numel()}, mismatching decided {int(bad.sum())}, equal {int((out.samples.cpu()[:, 0] == g['frames'][:, f, 0]).sum())}")
        nxt = g["frames"][:, f]
        run_ids = torch.cat([nxt, torch.zeros(B, 1, dtype=torch.long)], dim=1).unsqueeze(1)
        run_mask = torch.zeros(B, 1, 33, dtype=torch.int32)
        run_mask[:, :, :32] = 1
    model._drop_engine()
    del model
