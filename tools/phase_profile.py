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
    clocks = (C.c_uint64 * (16 * nph))()
    info = (C.c_int32 * (4 * nph))()
    for _ in range(2):
        e.call(e.lib.csm_debug_profile_frame, a.batch, clocks, info, e._stream())
    for cta, label in ((0, "first CTA"), (1, "last CTA")):
        t = [list(clocks[(cta * nph + ph) * 8:(cta * nph + ph) * 8 + 8]) for ph in range(nph)]
        total_cyc = t[nph - 1][2] - t[0][1]
        mhz = total_cyc / (ms / n * 1000.0)
        print(f"{label}: frame = {total_cyc} cycles (~{mhz:.0f} MHz if the profiled frame took the same time), {nph} phases")
        keys = ("body", "sync", "arrive", "poll", "bsync")
        sub = {k: defaultdict(float) for k in ("stage", "mma", "msync", "epi")}
        acc = {k: defaultdict(float) for k in keys}
        cnt = defaultdict(int)
        for ph in range(nph):
            ty, ep, stack, am = info[4 * ph], info[4 * ph + 1], info[4 * ph + 2], info[4 * ph + 3]
            kind = TYPES[ty] if ty != 1 else ("dec " if stack else "bb  ") + EPI[ep] + {3: " [K-stream]", 4: " [+attn]"}.get(am, "")
            cnt[kind] += 1
            acc["body"][kind] += t[ph][2] - t[ph][1]
            if ty == 1 and t[ph][6]:
                s4 = t[ph][4] or t[ph][1]
                sub["stage"][kind] += s4 - t[ph][1]
                sub["mma"][kind] += t[ph][5] - s4
                sub["msync"][kind] += t[ph][6] - t[ph][5]
                sub["epi"][kind] += t[ph][2] - t[ph][6]
            if ph + 1 < nph:
                acc["arrive"][kind] += t[ph][3] - t[ph][2]            # CTA sync + fence + atomic
                acc["poll"][kind] += t[ph + 1][0] - t[ph][3]          # waiting for the other CTAs
                acc["bsync"][kind] += t[ph + 1][1] - t[ph + 1][0]     # CTA sync + descriptor after the barrier
        print(f"{'phase kind':30s} {'n':>4s} {'body us':>8s} {'arrive':>7s} {'poll':>7s} {'bsync':>7s} {'total ms':>9s} {'share':>6s}  | {'stage':>6s} {'mma':>6s} {'msync':>6s} {'epi':>6s}")
        for k in sorted(cnt, key=lambda k: -sum(acc[x][k] for x in keys)):
            tot = sum(acc[x][k] for x in keys)
            c = cnt[k]
            print(f"{k:30s} {c:4d} {acc['body'][k] / c / mhz:8.2f} {acc['arrive'][k] / c / mhz:7.2f} "
                  f"{acc['poll'][k] / c / mhz:7.2f} {acc['bsync'][k] / c / mhz:7.2f} {tot / mhz / 1000:9.3f} {100 * tot / total_cyc:5.1f}%"
                  + (f"  | {sub['stage'][k] / c / mhz:6.2f} {sub['mma'][k] / c / mhz:6.2f} {sub['msync'][k] / c / mhz:6.2f} {sub['epi'][k] / c / mhz:6.2f}" if k in sub["mma"] else ""))


if __name__ == "__main__":
    main()
