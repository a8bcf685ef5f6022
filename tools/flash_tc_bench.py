"""Time and check the tcgen05 flash-attention forward (csrc/csm_flash_tc.cu) on its own: csm-1b backbone shapes
(32 query heads, 8 kv heads, head dim 64), one sequence of --seq tokens.  The check is torch SDPA in fp32 on the same
bf16 inputs (a checker, not a product path)."""
import argparse
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from csm_hf_b200 import native  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seq", type=int, default=4096)
ap.add_argument("--nseq", type=int, default=1)
ap.add_argument("--heads", type=int, default=32)
ap.add_argument("--kv", type=int, default=8)
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
lib = native.load()
lib.csm_flash_tc_launch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]
lib.csm_flash_tc_launch.restype = C.c_int
dev = torch.device("cuda", 0)
S, H, KV, HD = a.seq, a.heads, a.kv, 64
W = (H + 2 * KV) * HD
g = torch.Generator(device="cpu").manual_seed(0)
qkv = (torch.randn(a.nseq * S, W, generator=g) * 0.5).to(torch.bfloat16).to(dev)
out = torch.zeros(a.nseq * S, H * HD, dtype=torch.bfloat16, device=dev)
lse = torch.zeros(a.nseq * S, H, dtype=torch.float32, device=dev)
scale = HD ** -0.5
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def run():
    rc = lib.csm_flash_tc_launch(qkv.data_ptr(), S, a.nseq, H, KV, C.c_float(scale), None, out.data_ptr(), lse.data_ptr(), st)
    assert rc == 0, rc


run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
flops = 4 * a.nseq * H * HD * S * S / 2
print(f"flash_tc S={S} nseq={a.nseq}: {ms:.3f} ms per launch, {flops / ms / 1e9:.0f} TFLOP/s (causal)")
# check (first sequence, fp32 reference)
q = qkv[:S, : H * HD].view(S, H, HD).transpose(0, 1).float()
k = qkv[:S, H * HD: (H + KV) * HD].view(S, KV, HD).transpose(0, 1).float().repeat_interleave(H // KV, dim=0)
v = qkv[:S, (H + KV) * HD:].view(S, KV, HD).transpose(0, 1).float().repeat_interleave(H // KV, dim=0)
ref = torch.nn.functional.scaled_dot_product_attention(q[None], k[None], v[None], is_causal=True)[0]
got = out[:S].view(S, H, HD).transpose(0, 1).float()
err = (got - ref).abs().max().item()
print(f"max |err| vs fp32 SDPA: {err:.5f} (output scale {ref.abs().max().item():.3f})")
s = (q @ k.transpose(1, 2)) * scale
s = s.masked_fill(torch.ones(S, S, device=dev, dtype=torch.bool).triu(1), float("-inf"))
print(f"max |lse err|: {(torch.logsumexp(s, dim=-1).transpose(0, 1) - lse[:S]).abs().max().item():.5f}")
