"""Training-step benchmark (BASELINE config #5: csm-1b fwd+bwd, seq_len 4096, 1/16 decoder amortisation, batch 8 over
8 GPUs = one sequence per GPU): ms per step, tokens/s and the achieved fraction of the measured bf16 tensor peak.

  python tools/train_bench.py [--seq 4096] [--batch 1] [--steps 5] [--warmup 2] [--tiny]
  torchrun --nproc-per-node N tools/train_bench.py ...    data parallel: NCCL all-reduce of the gradients after backward

Synthetic batch (csm_hf_b200.synthetic.make_training_batch), random-init weights.  One step = model(ids, mask, labels)
+ loss.backward() (+ gradient all-reduce when N > 1): every kernel of it is in csrc/csm_train.cu / csm_gemm.cu.
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from csm_hf_b200.config import CSMConfig, tiny_config  # noqa: E402
from csm_hf_b200.modeling import CSMModel  # noqa: E402
from csm_hf_b200.synthetic import make_state_dict, make_training_batch  # noqa: E402


def flops_forward(cfg, B, S, F):
    b, d = cfg.backbone_config, cfg.decoder_config
    V = cfg.audio_vocab_size

    def stack(c):
        hd = c.hidden_size // c.num_attention_heads
        return c.num_hidden_layers * (c.hidden_size * (c.num_attention_heads + 2 * c.num_key_value_heads) * hd
                                      + c.hidden_size * c.hidden_size + 3 * c.hidden_size * c.intermediate_size)

    R, Rd = B * S, F * 33
    gemm = 2 * R * (stack(b) + V * b.hidden_size) + 2 * Rd * (stack(d) + b.hidden_size * d.hidden_size) \
        + 2 * F * 31 * d.hidden_size * V
    attn = 4 * B * b.num_attention_heads * (b.hidden_size // b.num_attention_heads) * (S * S / 2) * b.num_hidden_layers \
        + 4 * F * d.num_attention_heads * (d.hidden_size // d.num_attention_heads) * (33 * 33 / 2) * d.num_hidden_layers
    return gemm + attn


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seq", type=int, default=4096)
    ap.add_argument("--batch", type=int, default=1, help="sequences per GPU")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--ratio", type=int, default=16)
    ap.add_argument("--tiny", action="store_true")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        from csm_hf_b200.dist import allreduce_gradients
        from csm_hf_b200.training import enable_data_parallel
    cfg = tiny_config() if a.tiny else CSMConfig()
    model = CSMModel(cfg, make_state_dict(cfg, seed=0, dtype=torch.bfloat16), device=dev)
    model.requires_grad_(True)
    if world > 1 and os.environ.get("CSM_DDP_OVERLAP", "1") != "0":
        enable_data_parallel(model)      # gradient all-reduce overlapped with the backward (training.py)
    ids, mask, labels = make_training_batch(cfg, a.batch, a.seq, seed=100 + rank, text_frames=16, amortization_ratio=a.ratio)
    ids, mask, labels = ids.to(dev), mask.to(dev), labels.to(dev)
    F = int((labels[:, :, :32] != -100).all(dim=2).sum())

    def step():
        model.zero_grad(set_to_none=True)
        out = model(input_ids=ids, attention_mask=mask, labels=labels)
        out.loss.backward()
        if world > 1:
            allreduce_gradients(model)
        return out

    for _ in range(a.warmup):
        out = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(a.steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    peak = 1400.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        peak = float(json.load(open(p)).get("bf16_tflops_sustained", peak))
    fl = 3 * flops_forward(cfg, a.batch, a.seq, F)
    if rank == 0:
        print(json.dumps({
            "metric": "training_tokens_per_s", "value": world * a.batch * a.seq / (ms / 1000.0), "unit": "tokens/s",
            "n_gpus": world, "ms_per_step": ms, "steps": a.steps, "warmup": a.warmup, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"csm-{'tiny' if a.tiny else '1b'} fwd+bwd, seq_len {a.seq}, {a.batch} sequence(s) per GPU, "
                                   f"1/{a.ratio} decoder amortisation ({F} frames)", "wall_s": time.time() - t0},
            "loss": float(out.loss.detach()), "backbone_loss": float(out.backbone_loss), "decoder_loss": float(out.decoder_loss),
            "roofline": {"bound": "tensor", "achieved": fl / (ms / 1000.0) / 1e12, "peak": peak, "unit": "TFLOP/s",
                         "frac": fl / (ms / 1000.0) / 1e12 / peak, "algorithmic_flops_per_step": fl},
            "gpu_launches_per_step": model._train_engine.launches() // (a.steps + a.warmup),
            "max_memory_gb": torch.cuda.max_memory_allocated(dev) / 2**30,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
