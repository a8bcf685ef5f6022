"""Mint tests/golden/sample_topk.pt by running the UNMODIFIED reference sampler on CPU.

Build-container only (needs /root/reference).  Usage:  python oracle/make_golden_sampling.py

For a few fixed logits rows (fp32 values that are exactly representable in bf16, so the CUDA sampler sees the same
numbers) it calls the reference's `sample_topk` (modeling_csm.py:179-189) n times and stores the histogram of the
returned ids.  tests/test_gpu_parity.py::test_topk_sampling_matches_reference_histograms compares the CUDA sampler's
histogram with it (two-sample test): the draw streams differ (torch generator vs counter-based hash), the
distributions must not.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
import modeling_csm as M  # noqa: E402

CASES = [  # (name, V, logit std, topk, temperature, row seed)
    ("v2051_k50_t0.8", 2051, 2.0, 50, 0.8, 11),
    ("v2051_k5_t1.3", 2051, 1.0, 5, 1.3, 12),
    ("v67_k10_t1.0_ties", 67, 1.5, 10, 1.0, 13),
]
N = 20000


def main():
    torch.manual_seed(2026)
    out = {"n": N, "cases": {}}
    for name, V, std, k, T, seed in CASES:
        g = torch.Generator().manual_seed(seed)
        row = (torch.randn(V, generator=g) * std).to(torch.bfloat16).float()
        if "ties" in name:                      # the k-th value appears three times: the reference keeps all of them
            kth = torch.topk(row, k).values[-1]
            others = torch.nonzero(row < kth).flatten()[:2]
            row[others] = kth
        ids = M.sample_topk(row.unsqueeze(0).repeat(N, 1), k, T).flatten().long()
        out["cases"][name] = {"row": row.to(torch.bfloat16), "topk": k, "temperature": T,
                              "counts": torch.bincount(ids, minlength=V).to(torch.int32)}
        kept = int((torch.bincount(ids, minlength=V) > 0).sum())
        print(f"{name}: {kept} distinct ids drawn by the reference")
    path = os.path.join(ROOT, "tests", "golden", "sample_topk.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
