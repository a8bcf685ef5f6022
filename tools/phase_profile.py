"""Per-phase timing of one decode frame from inside the persistent kernel.

usage: python tools/phase_profile.py [--batch B] [--ctx T] [--tiny]
Prints, per phase kind (clock64 stamps of the first / last CTA): count, mean body time and its split
(wait for the input tags, rest of staging, wait for the first weight chunk, MMA, CTA sync, epilogue),
and from the %globaltimer stamps of every CTA: the spread between the first and the last CTA finishing
a phase, and the mean phase period.
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
    ap.add_argument("--both", action="store_true", help="also print the last CTA's table")
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
    nph, G = e.info(2), e.info(1)
    clocks = (C.c_uint64 * ((32 + G) * nph))()
    info = (C.c_int32 * (4 * nph))()
    for _ in range(2):
        e.call(e.lib.csm_debug_profile_frame, a.batch, clocks, info, e._stream())

    def kind_of(ph):
        ty, ep, stack, am = info[4 * ph], info[4 * ph + 1], info[4 * ph + 2], info[4 * ph + 3]
        return TYPES[ty] if ty != 1 else ("dec " if stack else "bb  ") + EPI[ep] + {3: " [K-stream]", 4: " [+attn]"}.get(am, "")

    for cta, label in ((0, "first CTA"), (1, "last CTA")):
        if cta == 1 and not a.both:
            break
        t = [list(clocks[(cta * nph + ph) * 16:(cta * nph + ph) * 16 + 16]) for ph in range(nph)]
        total_cyc = t[nph - 1][2] - t[0][1]
        mhz = total_cyc / (ms / n * 1000.0)
        print(f"{label}: frame = {total_cyc} cycles (~{mhz:.0f} MHz if the profiled frame took the same time), {nph} phases")
        cols = ("body", "sync", "poll", "stage", "wwait", "mma", "msync", "epi", "iters")
        acc = {k: defaultdict(float) for k in cols}
        cnt = defaultdict(int)
        for ph in range(nph):
            k = kind_of(ph)
            cnt[k] += 1
            acc["body"][k] += t[ph][2] - t[ph][1]
            if ph + 1 < nph:
                acc["sync"][k] += t[ph + 1][1] - t[ph][2]
            if info[4 * ph] == 1 and t[ph][6]:
                s4 = t[ph][4] or t[ph][1]
                s8 = t[ph][8] or t[ph][1]
                s7 = t[ph][7] or s4
                acc["poll"][k] += max(0, s8 - t[ph][1])
                acc["stage"][k] += max(0, s4 - s8)
                acc["wwait"][k] += max(0, s7 - s4)
                acc["mma"][k] += t[ph][5] - max(s7, s4)
                acc["msync"][k] += t[ph][6] - t[ph][5]
                acc["epi"][k] += t[ph][2] - t[ph][6]
                acc["iters"][k] += t[ph][9]
        print(f"{'phase kind':30s} {'n':>4s} {'body us':>8s} {'sync':>6s} {'total ms':>9s} {'share':>6s}  | {'poll':>6s} {'stage':>6s} {'wwait':>6s} {'mma':>6s} {'msync':>6s} {'epi':>6s} {'iters':>6s}")
        for k in sorted(cnt, key=lambda k: -(acc["body"][k] + acc["sync"][k])):
            tot = acc["body"][k] + acc["sync"][k]
            c = cnt[k]
            line = (f"{k:30s} {c:4d} {acc['body'][k] / c / mhz:8.2f} {acc['sync'][k] / c / mhz:6.2f} {tot / mhz / 1000:9.3f} "
                    f"{100 * tot / total_cyc:5.1f}%")
            if acc["mma"][k]:
                line += "  | " + " ".join(f"{acc[x][k] / c / mhz:6.2f}" for x in ("poll", "stage", "wwait", "mma", "msync", "epi"))
                line += f" {acc['iters'][k] / c:6.1f}"
            print(line)
    # skew between CTAs: globaltimer at the end of every phase of every CTA
    base = 32 * nph
    ends = [[clocks[base + c * nph + ph] for c in range(G)] for ph in range(nph)]
    spread, period = defaultdict(float), defaultdict(float)
    cnt = defaultdict(int)
    slow = defaultdict(int)
    for ph in range(1, nph - 1):
        k = kind_of(ph)
        cnt[k] += 1
        spread[k] += max(ends[ph]) - min(ends[ph])
        period[k] += max(ends[ph]) - max(ends[ph - 1])
        slow[max(range(G), key=lambda c: ends[ph][c])] += 1
    print(f"{'phase kind':30s} {'n':>4s} {'spread us':>10s} {'period us':>10s}   (globaltimer over all {G} CTAs)")
    for k in sorted(cnt, key=lambda k: -period[k]):
        print(f"{k:30s} {cnt[k]:4d} {spread[k] / cnt[k] / 1000:10.2f} {period[k] / cnt[k] / 1000:10.2f}")
    worst = sorted(slow.items(), key=lambda kv: -kv[1])[:8]
    print("CTAs most often last to finish a phase:", ", ".join(f"cta{c}:{n}" for c, n in worst))


if __name__ == "__main__":
    main()
