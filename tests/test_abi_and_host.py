"""CPU-side checks: the C-ABI library loads and exports every symbol include/csm_b200.h declares;
host logic (config mirror, synthetic state_dict, RoPE tables, batch sharding over gloo)."""
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "csm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(csm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from csm_hf_b200 import build, native
    build.build()                      # nvcc cross-compiles without a GPU
    lib = native.load()
    names = header_functions()
    assert len(names) >= 13
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/csm_b200.h but not exported"
    assert set(native.EXPORTS) == set(names)


def test_engine_refuses_to_run_without_gpu():
    """No CPU fallback: constructing the engine off-GPU raises instead of degrading."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from csm_hf_b200.config import tiny_config
    from csm_hf_b200.modeling import CSMModel
    from csm_hf_b200.synthetic import make_context, make_state_dict
    cfg = tiny_config()
    with pytest.raises((RuntimeError, AssertionError)):
        m = CSMModel(cfg, make_state_dict(cfg), device="cpu")
        ids, mask = make_context(cfg, 1, 2)
        m.generate(ids, mask, max_new_frames=1, temperature=0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "csm_hf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_config_mirror_and_state_dict_keys():
    from csm_hf_b200.config import CSMConfig
    from csm_hf_b200.synthetic import state_dict_shapes
    cfg = CSMConfig()
    assert cfg.backbone_config.max_position_embeddings == 2048 and cfg.decoder_config.max_position_embeddings == 32
    shapes = state_dict_shapes(cfg)
    assert len(shapes) == 187                                   # the reference state_dict
    assert shapes["audio_head"] == (31, 1024, 2051)
    assert shapes["audio_embeddings.weight"] == (65632, 2048)
    n = sum(int(torch.tensor(s).prod()) for s in shapes.values())
    assert n == 1552791552                                      # SURVEY.md §6 parameter count
    rt = CSMConfig.from_dict(cfg.to_dict())
    assert rt.to_dict() == cfg.to_dict()


def test_rope_tables_match_oracle():
    from csm_hf_b200 import rope
    from csm_hf_b200.config import CSMConfig
    from oracle.csm_oracle import llama3_inv_freq, rope_cos_sin
    cfg = CSMConfig()
    for d, n in ((cfg.backbone_config, 2300), (cfg.decoder_config, 32)):
        cos, sin = rope.tables(d.head_dim, d.rope_theta, d.rope_scaling, n)
        f = llama3_inv_freq(d.head_dim, d.rope_theta, d.rope_scaling)
        c2, s2 = rope_cos_sin(f, torch.arange(n), torch.bfloat16)
        assert torch.equal(cos, c2[:, : d.head_dim // 2]) and torch.equal(sin, s2[:, : d.head_dim // 2])
    f = rope.inv_freq(64, 500000.0, cfg.backbone_config.rope_scaling)
    assert abs(float(f[0]) - 1.0) < 1e-7 and abs(float(f[-1]) - 9.418e-8) / 9.418e-8 < 1e-3


def test_shard_bounds_and_stop_rule():
    from csm_hf_b200.dist import shard_bounds, truncate_at_global_stop
    for B in (1, 7, 8, 64):
        for W in (1, 2, 4, 8):
            spans = [shard_bounds(B, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    fr = torch.ones(3, 5, 32, dtype=torch.long)
    fr[:, 3] = 0
    fr[0, 1] = 0                       # one sequence all-zero is NOT a stop (modeling_csm.py:662)
    assert truncate_at_global_stop(fr).shape[1] == 3
    assert truncate_at_global_stop(torch.ones(2, 4, 32, dtype=torch.long)).shape[1] == 4


WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from csm_hf_b200.dist import generate_sharded, shard_bounds
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
class Fake:
    """Deterministic per-sequence frames, so the gathered result can be checked exactly.  generate() emits the first
    chunk, generate_more() continues the same sequence of frames (as the engine's csm_generate_more does)."""
    device = "cpu"
    calls = 0
    def _frames(self, base, start, n):
        f = (torch.arange(start, start + n).view(1, -1, 1) * 32 + torch.arange(32).view(1, 1, -1) + 1)
        out = base + f
        out[:, max(0, 4 - start):] = 0                                   # every sequence goes silent at frame 4
        return out
    def generate(self, ids, mask, max_new_frames, temperature, topk, use_cache, stop_on_all_zeros, reserve_frames=0):
        assert stop_on_all_zeros is False
        self.base = ids[:, 0, 0:1].unsqueeze(1) * 1000                   # [b,1,1]
        self.done = max_new_frames
        Fake.calls += 1
        return self._frames(self.base, 0, max_new_frames)
    def generate_more(self, batch, n, stop_on_all_zeros=True):
        assert stop_on_all_zeros is False and batch == self.base.shape[0]
        out = self._frames(self.base, self.done, n)
        self.done += n
        Fake.calls += 1
        return out
B = 5
ids = torch.arange(B).view(B, 1, 1).repeat(1, 3, 33)
ref = Fake()
want = ref._frames(ids[:, 0, 0:1].unsqueeze(1) * 1000, 0, 12)
# one all-gather at the end when nothing can stop the loop early
m = Fake(); Fake.calls = 0
keep = generate_sharded(m, ids, None, max_new_frames=12, temperature=0, stop_on_all_zeros=False)
assert keep.shape == (B, 12, 32) and torch.equal(keep, want) and Fake.calls == 1
# chunks of 2 frames: the all-zero frame 4 is seen in the third chunk, the remaining 6 frames are never generated
m = Fake(); Fake.calls = 0
full = generate_sharded(m, ids, None, max_new_frames=12, temperature=0, stop_on_all_zeros=True, stop_check_every=2)
assert torch.equal(full, want[:, :4]), (full.shape,)
assert Fake.calls == 3 and m.done == 6
assert getattr(m, "seq_base", None) == 0                                 # restored after the call
# stop rule never fires: the chunked path returns every frame
class Loud(Fake):
    def _frames(self, base, start, n):
        return base + (torch.arange(start, start + n).view(1, -1, 1) * 32 + torch.arange(32).view(1, 1, -1) + 1)
loud = generate_sharded(Loud(), ids, None, max_new_frames=7, temperature=0, stop_on_all_zeros=True, stop_check_every=3)
assert loud.shape == (B, 7, 32)
dist.barrier(); dist.destroy_process_group()
print("ok")
'''


def test_sharded_generate_two_ranks_gloo(tmp_path):
    """world_size-2 run of the sharding + all-gather + stop-flag exchange between chunks on CPU (gloo)."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and "ok" in o, o


def test_ctypes_structs_match_the_header(tmp_path):
    """The ctypes mirror of CsmLlamaShape / CsmShapes / CsmWeights has the C compiler's size and field offsets."""
    import ctypes as C
    from csm_hf_b200 import native
    src = tmp_path / "layout.c"
    src.write_text(r"""
#include <stdio.h>
#include <stddef.h>
#include "csm_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\n", sizeof(CsmLlamaShape), offsetof(CsmLlamaShape, eps), offsetof(CsmLlamaShape, rope_cos),
         offsetof(CsmLlamaShape, rope_sin), offsetof(CsmLlamaShape, n_pos));
  printf("%zu %zu %zu\n", sizeof(CsmShapes), offsetof(CsmShapes, backbone), offsetof(CsmShapes, decoder));
  printf("%zu %zu %zu %zu\n", sizeof(CsmWeights), offsetof(CsmWeights, audio_head), offsetof(CsmWeights, backbone_layers),
         offsetof(CsmWeights, decoder_layers));
  return 0;
}
""")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    got = [int(x) for x in out]
    L, S, W = native.LlamaShape, native.Shapes, native.Weights
    want = [C.sizeof(L), L.eps.offset, L.rope_cos.offset, L.rope_sin.offset, L.n_pos.offset,
            C.sizeof(S), S.backbone.offset, S.decoder.offset,
            C.sizeof(W), W.audio_head.offset, W.backbone_layers.offset, W.decoder_layers.offset]
    assert got == want


def test_reference_arm_prints_the_contract_line():
    """bench.py --impl reference at BASELINE.json configs[0] size.  BENCH_CPU_PORT=1 times the oracle port (seconds);
    without it the arm drives the staged reference itself (oracle/_ref, minutes at csm-1b: run on the GPU box)."""
    import json
    env = dict(os.environ, BENCH_CPU_PORT="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ctx", "16", "--frames", "8", "--cpu-sample-frames", "1"], capture_output=True, text=True,
                         timeout=600, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "audio_frames_per_s" and line["unit"] == "frames/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["projected"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["projected"] is True
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0


def test_staged_reference_is_what_the_harness_drives():
    """oracle/stage_ref.py copies the reference's files UNCHANGED under oracle/_ref (git-ignored); ref_harness falls back
    to that copy where /root/reference does not exist (the GPU box)."""
    from oracle import ref_harness, stage_ref
    if not stage_ref.stage():
        pytest.skip("no reference available in this environment")
    staged = os.path.join(ROOT, "oracle", "_ref", "modeling_csm.py")
    assert os.path.isfile(staged) and ref_harness.reference_available()
    src = os.path.join(stage_ref.REF_SRC, "modeling_csm.py")
    if os.path.isfile(src):
        assert open(src, "rb").read() == open(staged, "rb").read()
    ign = open(os.path.join(ROOT, ".gitignore")).read()
    assert "oracle/_ref/" in ign


def _nvcc():
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.isfile(nvcc):
        pytest.skip("nvcc not available")
    return nvcc


def test_product_library_is_blackwell_native():
    """The shipped library runs its prefill projections on tcgen05 / TMEM / TMA: the SASS of libcsm_b200.so contains
    the instructions (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, cp.async.bulk.tensor -> UTMALDG,
    cp.async.bulk -> UBLKCP), and it links no cuBLAS."""
    _nvcc()
    from csm_hf_b200 import build
    lib = build.build()
    sass = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "UBLKCP", "HMMA.16816"):
        assert mnemonic in sass, mnemonic
    ldd = subprocess.run(["ldd", lib], capture_output=True, text=True).stdout
    assert "cublas" not in ldd.lower()


def test_tcgen05_prototype_still_compiles(tmp_path):
    """tools/micro/umma_gemm.cu: the stand-alone self-checking GEMM the product kernel (csm_gemm.cu) grew out of
    (run on a B200: profiles/r02_umma_gemm_selfcheck.txt)."""
    nvcc = _nvcc()
    exe = tmp_path / "umma_gemm"
    g = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-o", str(exe),
                        os.path.join(ROOT, "tools", "micro", "umma_gemm.cu")], capture_output=True, text=True)
    assert g.returncode == 0, g.stderr[-2000:]


GRAD_WORKER = r'''
import sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from csm_hf_b200.dist import allreduce_gradients
rank = int(sys.argv[3])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=rank, world_size=2)
torch.manual_seed(0)
m = torch.nn.ModuleDict({"a": torch.nn.Linear(64, 48, bias=False), "b": torch.nn.Linear(48, 8, bias=False),
                         "c": torch.nn.Embedding(100, 64)}).to(torch.bfloat16)
both = []
for r in range(2):
    g = torch.Generator().manual_seed(10 + r)
    both.append({k: torch.randn(p.shape, generator=g).to(torch.bfloat16) for k, p in m.named_parameters()})
for k, p in m.named_parameters():
    p.grad = both[rank][k].clone()
n = allreduce_gradients(m, bucket_bytes=8000)            # several buckets
assert n >= 2, n
for k, p in m.named_parameters():
    want = ((both[0][k].float() + both[1][k].float()) / 2).to(torch.bfloat16)
    assert torch.equal(p.grad, want), k
dist.barrier(); dist.destroy_process_group()
print("ok")
'''


def test_gradient_allreduce_two_ranks_gloo(tmp_path):
    """Data-parallel training plumbing: bucketed average of the parameter gradients over 2 ranks."""
    script = tmp_path / "gworker.py"
    script.write_text(GRAD_WORKER)
    port = str(31500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and "ok" in o, o
