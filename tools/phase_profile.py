"""Per-phase timing of one decode frame from inside the persistent kernel (CTA 0 clock64 stamps).

usage: python tools/phase_profile.py [--batch B] [--ctx T] [--tiny]
Prints, per phase kind: count, mean body time, mean wait (barrier + idle before the phase).
"""
import argparse
import ctypes as C
import os
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from csm_hf_b200.config import CSMConfig, tiny_config  # noqa: E402
from csm_hf_b200.modeling import CSMModel  # noqa: E402
from csm_hf_b200.synthetic import make_context, make_state_dict  # noqa: E402

TYPES = {0: "embed", 1: "gemv", 2: "attn_bb", 3: "attn_dec", 4: "finish"}
EPI = {0: "store(proj)", 1: "resid", 2: "swiglu(gate/up)", 3: "qkv+rope", 4: "head+argmax"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--ctx", type=int, default=2048)
    ap.add_argument("--tiny", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    cfg = tiny_config() if a.tiny else CSMConfig()
    sd = make_state_dict(cfg, seed=0, dtype=torch.bfloat16)
    model = CSMModel(cfg, sd, device=dev, max_batch=a.batch, max_ctx=a.ctx + 64)
    ids, mask = make_context(cfg, a.batch, a.ctx)
    model.generate(ids.to(dev), mask.to(dev), max_new_frames=8, temperature=0, stop_on_all_zeros=False)
    ms, n = model.last_decode_ms()
    print(f"decode frame (events): {ms / n:.3f} ms")
    e = model.engine()
    nph = e.info(2)
    clocks = (C.c_uint64 * (2 * nph))()
    info = (C.c_int32 * (4 * nph))()
    for _ in range(2):
        e.call(e.lib.csm_debug_profile_frame, a.batch, clocks, info, e._stream())
    t = list(clocks)
    total_cyc = t[2 * nph - 1] - t[0]
    mhz = total_cyc / (ms / n * 1000.0)
    print(f"frame = {total_cyc} cycles of CTA 0  (~{mhz:.0f} MHz if the profiled frame took the same time)")
    body, wait, cnt = defaultdict(float), defaultdict(float), defaultdict(int)
    for ph in range(nph):
        ty, ep, stack, am = info[4 * ph], info[4 * ph + 1], info[4 * ph + 2], info[4 * ph + 3]
        kind = TYPES[ty] if ty != 1 else ("dec " if stack else "bb  ") + EPI[ep] + (" [K-stream]" if am == 3 else "")
        b = t[2 * ph + 1] - t[2 * ph]
        w = t[2 * ph] - t[2 * ph - 1] if ph > 0 else 0
        body[kind] += b
        wait[kind] += w
        cnt[kind] += 1
    print(f"{'phase kind':34s} {'n':>4s} {'body us':>9s} {'wait us':>9s} {'total ms':>9s} {'share':>6s}")
    for k in sorted(cnt, key=lambda k: -(body[k] + wait[k])):
        tot = body[k] + wait[k]
        print(f"{k:34s} {cnt[k]:4d} {body[k] / cnt[k] / mhz:9.2f} {wait[k] / cnt[k] / mhz:9.2f} "
              f"{tot / mhz / 1000:9.3f} {100 * tot / total_cyc:5.1f}%")


if __name__ == "__main__":
    main()
